/* wavefront.cuh - STAGE 1 placeholder (replaced by the real wavefront). */
static void free_pool(b200_ctx *) {}
static bool svm_validate(const uint32_t *, size_t, std::string &) { return true; }
static int check_scope(b200_ctx *) { return B200_OK; }
extern "C" {
int b200_render(b200_ctx *ctx, const b200_work_tile *, volatile const int *) { return fail(ctx, B200_ERR_NOT_READY, "render not built yet"); }
int b200_film_convert(b200_ctx *ctx, uint64_t, uint64_t, int, float, int, int, int, int, int, int) { return fail(ctx, B200_ERR_NOT_READY, "nyi"); }
int b200_film_reduce(b200_ctx **, int, const uint64_t *, size_t) { return B200_ERR_NOT_READY; }
}
