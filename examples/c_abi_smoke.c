/* Plain-C caller of libmrs_b200.so: what a cgo / JNI / N-API binding would do at load time.
 * No CUDA headers, no Python.  Build and run (no GPU needed for these calls):
 *   gcc -std=c99 -Iinclude examples/c_abi_smoke.c -o /tmp/c_abi_smoke \
 *       -Lmrs-gym_b200/mrsgym_b200 -lmrs_b200 -Wl,-rpath,$PWD/mrs-gym_b200/mrsgym_b200 && /tmp/c_abi_smoke
 * With a GPU the same program would go on to cudaMalloc the buffers of MrsBuffers and call mrs_step. */
#include <stdio.h>
#include <string.h>

#include "mrs_b200.h"

int main(void) {
    MrsConfig cfg;
    if (mrs_abi_version() != MRS_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
    if (mrs_sizeof_config() != sizeof(MrsConfig) || mrs_sizeof_buffers() != sizeof(MrsBuffers)) {
        fprintf(stderr, "struct layout mismatch\n");
        return 1;
    }
    if (mrs_default_config(&cfg) != MRS_OK) return 1;
    cfg.E = 65536; cfg.N = 8; cfg.K = 3; cfg.L = 16;
    cfg.action_type = MRS_SET_SPEEDS;
    cfg.comm_range = 2.0f;
    printf("abi %d, sizeof(MrsConfig) %zu, action dim %d, state dim %d, scratch planes %d, baked %d\n", mrs_abi_version(),
           sizeof(MrsConfig), mrs_action_dim(cfg.action_type), mrs_state_dim(cfg.state_layout),
           mrs_scratch_planes(cfg.E, cfg.N), mrs_config_is_baked(&cfg));
    printf("hover: mass %.4f kg, kf %.3e, dt %.3f s; error text for -1: %s\n", cfg.quad.mass, cfg.quad.kf, cfg.dt,
           mrs_strerror(MRS_ERR_ARG));
    /* argument validation happens before any CUDA call */
    if (mrs_step(&cfg, NULL, NULL, 0, 0, NULL) != MRS_ERR_ARG) return 1;
    return mrs_config_is_baked(&cfg) == 1 ? 0 : 1;
}
