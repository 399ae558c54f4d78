// tcgen05 engine (placeholder).
#include "score.h"
namespace dpb {
int tc_prepare(dpb_score* h, const dpb_score_weights*) { h->tc_ready = false; return DPB_OK; }
void tc_release(dpb_score*) {}
int tc_launch(dpb_score*, const TcJob&, cudaStream_t) { return fail(DPB_EUNSUPPORTED, "tcgen05 engine not built"); }
}  // namespace dpb
