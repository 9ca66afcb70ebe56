// encoder_api.cu — C ABI of the session encoder and the item vote (filled in below as the kernels land).
#include "../../include/sss_b200.h"
#include "common.cuh"

using namespace sss;

extern "C" int sss_item_vote(const float*, const int64_t*, int64_t, int, const int64_t*, const int64_t*, int64_t, int,
                             int64_t*, float*, int, void*) {
  set_error("sss_item_vote: not built yet");
  return 1;
}
extern "C" int sss_encoder_create(sss_encoder_t**, int, const sss_encoder_shape_t*) {
  set_error("sss_encoder_create: not built yet");
  return 1;
}
extern "C" int sss_encoder_destroy(sss_encoder_t*) { return 0; }
extern "C" int sss_encoder_set_param(sss_encoder_t*, const char*, const float*, int64_t, int, void*) {
  set_error("sss_encoder_set_param: not built yet");
  return 1;
}
extern "C" int sss_encoder_forward(sss_encoder_t*, const sss_graph_batch_t*, float*, int32_t*, void*) {
  set_error("sss_encoder_forward: not built yet");
  return 1;
}
extern "C" int sss_binarize_head(const float*, const float*, const float*, int64_t, int, int, float*, int, void*) {
  set_error("sss_binarize_head: not built yet");
  return 1;
}
