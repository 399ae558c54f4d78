// tcgen05 pose-blend engine for LBS (placeholder until the tensor-core kernel lands).
#include "lbs.h"

namespace dpb {
int lbs_tc_prepare(dpb_lbs* h, const dpb_body_tensors* m) { (void)m; h->tc_ready = false; return DPB_OK; }
void lbs_tc_release(dpb_lbs* h) { (void)h; }
int lbs_tc_vertices(dpb_lbs* h, const float*, const float*, const LbsWs&, float*, int64_t, cudaStream_t) {
  (void)h;
  return fail(DPB_EUNSUPPORTED, "LBS tensor-core engine not built");
}
}  // namespace dpb

