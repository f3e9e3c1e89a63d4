/* Link-time stubs for symbols the reference host code names but the hot path
 * never reaches (SURVEY.md §8c "Link-time gaps and stubs"):
 *   - OpenGL / GLEW entry points used only by Device::draw_pixels
 *     (intern/cycles/device/device.cpp) - libGL is absent from the snapshot;
 *   - OIIOImageLoader (intern/cycles/render/image_oiio.cpp) - libOpenImageIO.a is
 *     absent; no config uses image textures, so loading always fails;
 *   - numaAPI_* - numa.h is absent; report "no NUMA", as the reference's own
 *     numaapi_stub.c does on platforms without libnuma.
 * TEST INFRASTRUCTURE ONLY: part of oracle/_ref/libcycles_ref.so. */

#include <stddef.h>

#include "numaapi.h"
#include "render/image_oiio.h"

extern "C" {

#define GLEW_PTR(name) void *name = NULL;
GLEW_PTR(__glewUseProgram)
GLEW_PTR(__glewGetShaderInfoLog)
GLEW_PTR(__glewBindBuffer)
GLEW_PTR(__glewVertexAttribPointer)
GLEW_PTR(__glewUnmapBuffer)
GLEW_PTR(__glewUniform2f)
GLEW_PTR(__glewUniform1i)
GLEW_PTR(__glewShaderSource)
GLEW_PTR(__glewMapBuffer)
GLEW_PTR(__glewLinkProgram)
GLEW_PTR(__glewGetUniformLocation)
GLEW_PTR(__glewGetShaderiv)
GLEW_PTR(__glewGetProgramiv)
GLEW_PTR(__glewGetAttribLocation)
GLEW_PTR(__glewGenVertexArrays)
GLEW_PTR(__glewGenBuffers)
GLEW_PTR(__glewEnableVertexAttribArray)
GLEW_PTR(__glewDeleteVertexArrays)
GLEW_PTR(__glewDeleteProgram)
GLEW_PTR(__glewDeleteBuffers)
GLEW_PTR(__glewCreateShader)
GLEW_PTR(__glewCreateProgram)
GLEW_PTR(__glewCompileShader)
GLEW_PTR(__glewBufferData)
GLEW_PTR(__glewBindVertexArray)
GLEW_PTR(__glewBindFragDataLocation)
GLEW_PTR(__glewAttachShader)
GLEW_PTR(__glewActiveTexture)

#define GL_FN(name) \
  void name(void) \
  { \
  }
GL_FN(glTexParameteri)
GL_FN(glTexImage2D)
GL_FN(glBindTexture)
GL_FN(glGetIntegerv)
GL_FN(glGenTextures)
GL_FN(glEnable)
GL_FN(glDrawArrays)
GL_FN(glDisable)
GL_FN(glDeleteTextures)
GL_FN(glBlendFunc)

NUMAAPI_Result numaAPI_Initialize(void)
{
  return NUMAAPI_NOT_AVAILABLE;
}
int numaAPI_GetNumNodes(void)
{
  return 0;
}
bool numaAPI_IsNodeAvailable(int)
{
  return false;
}
int numaAPI_GetNumNodeProcessors(int)
{
  return 0;
}
int numaAPI_GetNumCurrentNodesProcessors(void)
{
  return 0;
}
bool numaAPI_RunThreadOnNode(int)
{
  return false;
}

} /* extern "C" */

CCL_NAMESPACE_BEGIN

OIIOImageLoader::OIIOImageLoader(const string &filepath) : filepath(filepath)
{
}
OIIOImageLoader::~OIIOImageLoader()
{
}
bool OIIOImageLoader::load_metadata(ImageMetaData &)
{
  return false;
}
bool OIIOImageLoader::load_pixels(const ImageMetaData &, void *, const size_t, const bool)
{
  return false;
}
string OIIOImageLoader::name() const
{
  return filepath.string();
}
ustring OIIOImageLoader::osl_filepath() const
{
  return filepath;
}
bool OIIOImageLoader::equals(const ImageLoader &other) const
{
  return filepath == ((const OIIOImageLoader &)other).filepath;
}

CCL_NAMESPACE_END
