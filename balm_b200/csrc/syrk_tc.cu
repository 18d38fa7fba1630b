// syrk_tc.cu -- tcgen05 (5th-gen tensor core) path of the rank-3M symmetric update. Placeholder until the
// split-integer kernel lands: BALM_PREC_TENSOR contexts fail loudly instead of silently falling back.
#include "internal.cuh"

int tensor_syrk_init(balm_ctx *c) {
  balm_set_error("BALM_PREC_TENSOR: tcgen05 SYRK not built in this version");
  return BALM_ERR_UNSUPPORTED;
}
int launch_tensor_syrk(balm_ctx *c, int64_t rows, bool first_batch) {
  balm_set_error("BALM_PREC_TENSOR: tcgen05 SYRK not built in this version");
  return BALM_ERR_UNSUPPORTED;
}
void tensor_syrk_free(balm_ctx *c) {}
