// placeholder until the fused fitting step lands
#include "engine.cuh"
using namespace nphm;
extern "C" long long nphm_fit_workspace_bytes(const nphm_ensemble *, long long) { return 0; }
extern "C" int nphm_fit_identity_step(nphm_ensemble *, const float *, long long, float *, float *, float *,
                                      const nphm_fit_params *, int, float *, float *, void *, void *)
{ set_error("fit step not built"); return NPHM_ERR_UNSUPPORTED; }
