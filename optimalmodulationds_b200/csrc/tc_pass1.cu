// placeholder: the tcgen05 prefilter is added in a later commit; until then every mode resolves to fp32.
#include "internal.cuh"
int tc_build_images(dsmppi_ctx* c, const dsmppi_net*) { c->tc_blob = nullptr; return 0; }
int tc_set_obstacles(dsmppi_ctx*, cudaStream_t) { return 0; }
int tc_pass1(dsmppi_ctx*, const float*, int, int, uint32_t, int, cudaStream_t) {
  dsmppi_set_error("tensor-core pass 1 not built");
  return 3;
}
