/* placeholder until the tcgen05 kernels land (next commit) */
#include "gemm.cuh"
namespace psb {
void fc_forward_tf32(Ctx*, const FcFwdArgs&) { throw Error(PS_ERR_ARG, "PS_FC_TF32 not built"); }
void fc_dgrad_tf32(Ctx*, const FcDgradArgs&) { throw Error(PS_ERR_ARG, "PS_FC_TF32 not built"); }
void fc_wgrad_tf32(Ctx*, const FcWgradArgs&) { throw Error(PS_ERR_ARG, "PS_FC_TF32 not built"); }
}
