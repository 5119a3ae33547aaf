// hdp.cu — placeholder until the HDP kernels land (replaced later in this round).
#include "model.hpp"
using namespace bias;
namespace bias {
int ensure_stirling(bias_ctx *, int32_t) { return fail(BIAS_EINVAL, "HDP not built yet"); }
int hdp_create(bias_model *, const bias_hdp_params *) { return fail(BIAS_EINVAL, "HDP not built yet"); }
int hdp_sweep(bias_model *) { return fail(BIAS_EINVAL, "HDP not built yet"); }
int hdp_destroy(bias_model *) { return BIAS_OK; }
}
extern "C" {
int bias_model_get_hdp(bias_model *, bias_hdp_params *, double *) { return fail(BIAS_EINVAL, "HDP not built yet"); }
int bias_hdp_table_conditionals(bias_ctx *, int64_t, const int32_t *, const double *, const double *, int32_t, double *, int32_t *) { return fail(BIAS_EINVAL, "HDP not built yet"); }
int bias_stirling_rows(bias_ctx *, int32_t, double *) { return fail(BIAS_EINVAL, "HDP not built yet"); }
}
