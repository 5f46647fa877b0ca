// placeholder: tensor-core path (filled in next)
#include "tz_prednet.cuh"
namespace tz {
int tc_create(tz_prednet *h, const std::vector<std::vector<float>> &) { set_error("tensor-core path not built"); return TZ_ECUDA; }
void tc_destroy(tz_prednet *) {}
int tc_next(tz_prednet *, const float *, float *, int, cudaStream_t) { set_error("tensor-core path not built"); return TZ_ECUDA; }
}
