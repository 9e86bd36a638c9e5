// Entry points declared in include/picgpu.h whose kernels are not written yet.  They fail loudly.
#include "common.cuh"
using namespace picg;
extern "C" {
int picg_species_push_heavy(picg_species_t, picg_species_t, picg_species_t, double, int) { return set_error(PICG_ERR_STATE, "picg_species_push_heavy: not implemented yet"); }
int picg_mcc_create(picg_species_t, picg_species_t, picg_species_t, picg_world_t, const double*, const double*, int, double, picg_mcc_t*) { return set_error(PICG_ERR_STATE, "picg_mcc_create: not implemented yet"); }
int picg_mcc_destroy(picg_mcc_t) { return PICG_OK; }
int picg_mcc_apply(picg_mcc_t, double, picg_mcc_stats*) { return set_error(PICG_ERR_STATE, "picg_mcc_apply: not implemented yet"); }
int picg_mcc_set_wsv_max(picg_mcc_t, double) { return set_error(PICG_ERR_STATE, "picg_mcc_set_wsv_max: not implemented yet"); }
int picg_mcc_sigma(picg_mcc_t, int, const double*, double*, double*) { return set_error(PICG_ERR_STATE, "picg_mcc_sigma: not implemented yet"); }
int picg_source_create(picg_species_t, picg_world_t, double, double, double, int, picg_source_t*) { return set_error(PICG_ERR_STATE, "picg_source_create: not implemented yet"); }
int picg_source_destroy(picg_source_t) { return PICG_OK; }
int picg_source_sample(picg_source_t, size_t*) { return set_error(PICG_ERR_STATE, "picg_source_sample: not implemented yet"); }
}
