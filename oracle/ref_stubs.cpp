/* Link-time stubs for symbols the reference host code names but the hot path
 * never reaches (SURVEY.md §8c "Link-time gaps and stubs"):
 *   - OpenGL / GLEW entry points used only by Device::draw_pixels
 *     (intern/cycles/device/device.cpp) - libGL is absent from the snapshot;
 *   - OIIOImageLoader (intern/cycles/render/image_oiio.cpp) - libOpenImageIO.a is
 *     absent; the stand-in below reads this repo's raw test-image container instead;
 *   - numaAPI_* - numa.h is absent; report "no NUMA", as the reference's own
 *     numaapi_stub.c does on platforms without libnuma.
 * TEST INFRASTRUCTURE ONLY: part of oracle/_ref/libcycles_ref.so. */

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "numaapi.h"
#include "render/image_oiio.h"
#include "util/util_half.h"
#include <vector>

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

/* Stand-in for image_oiio.cpp.  Instead of the image formats OpenImageIO decodes it reads
 * this repo's own raw container (written by raytracingproject_b200/scenes.py:
 * write_b2im), so that the reference's ImageManager - colour-space detection, RGB(A)
 * widening, device_texture upload, TextureInfo - runs unmodified on test images:
 *   char magic[4] = "B2IM"; uint32 width, height, channels, kind;
 *   kind 0 = uint8, 1 = float32, 2 = float16, 3 = uint16
 *   then height rows, bottom row first (the order the kernel samples in), `channels`
 *   interleaved values per pixel.
 * load_metadata mirrors what the real loader derives from the file's ImageSpec
 * (image_oiio.cpp:43-102): 1-channel files stay single-channel textures, everything else
 * becomes a 4-channel texture of the file's storage type. */
struct B2imHeader {
  char magic[4];
  uint32_t width, height, channels, kind;
};

static bool b2im_header(const char *path, B2imHeader &h)
{
  FILE *f = fopen(path, "rb");
  if (!f)
    return false;
  const bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "B2IM", 4) == 0 &&
                  h.channels >= 1 && h.channels <= 4 && h.kind <= 3;
  fclose(f);
  return ok;
}

OIIOImageLoader::OIIOImageLoader(const string &filepath) : filepath(filepath)
{
}
OIIOImageLoader::~OIIOImageLoader()
{
}
bool OIIOImageLoader::load_metadata(ImageMetaData &metadata)
{
  B2imHeader h;
  if (!b2im_header(filepath.c_str(), h))
    return false;
  metadata.width = h.width;
  metadata.height = h.height;
  metadata.depth = 1;
  metadata.channels = (int)h.channels;
  const bool rgba = h.channels > 1;
  switch (h.kind) {
    case 0:
      metadata.type = rgba ? IMAGE_DATA_TYPE_BYTE4 : IMAGE_DATA_TYPE_BYTE;
      break;
    case 1:
      metadata.type = rgba ? IMAGE_DATA_TYPE_FLOAT4 : IMAGE_DATA_TYPE_FLOAT;
      break;
    case 2:
      metadata.type = rgba ? IMAGE_DATA_TYPE_HALF4 : IMAGE_DATA_TYPE_HALF;
      break;
    default:
      metadata.type = rgba ? IMAGE_DATA_TYPE_USHORT4 : IMAGE_DATA_TYPE_USHORT;
      break;
  }
  metadata.colorspace_file_format = "b2im";
  return true;
}
bool OIIOImageLoader::load_pixels(const ImageMetaData &metadata,
                                  void *pixels,
                                  const size_t pixels_size,
                                  const bool)
{
  B2imHeader h;
  if (!b2im_header(filepath.c_str(), h))
    return false;
  static const size_t value_bytes[4] = {1, 4, 2, 2};
  const size_t values = (size_t)h.width * h.height * h.channels;
  if (values != pixels_size || (size_t)metadata.width != h.width)
    return false;
  FILE *f = fopen(filepath.c_str(), "rb");
  if (!f)
    return false;
  fseek(f, (long)sizeof(B2imHeader), SEEK_SET);
  bool ok;
  if (h.kind == 2 &&
      (metadata.type == IMAGE_DATA_TYPE_FLOAT || metadata.type == IMAGE_DATA_TYPE_FLOAT4)) {
    /* a device without half images (DeviceInfo::has_half_images, image.cpp:354-362) gets
     * the file widened to float, as the real loader's typed read does */
    std::vector<half> tmp(values);
    ok = fread(tmp.data(), 2, values, f) == values;
    for (size_t i = 0; ok && i < values; i++)
      ((float *)pixels)[i] = half_to_float(tmp[i]);
  }
  else {
    ok = fread(pixels, value_bytes[h.kind], values, f) == values;
  }
  fclose(f);
  return ok;
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
