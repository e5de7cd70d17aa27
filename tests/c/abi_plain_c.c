/* include/rbp.h used from plain C11: the boundary is a C ABI (no C++ or torch types in any signature). */
#include <stdio.h>

#include "rbp.h"

int main(void) {
    rbp_hyper_t h;
    uint32_t counter[4] = {0, 0, 0, 0}, key[2] = {0, 0}, out[4];
    rbp_solver_t* s = NULL;
    int status;
    rbp_hyper_default(&h);
    rbp_philox4x32_10(counter, key, out);
    if (out[0] != 0x6627E8D5u || h.prune_warmup != 16384u) return 1;
    status = rbp_solver_create(RBP_GAME_KUHN, RBP_REGRET_FLOORED, RBP_WEIGHT_LINEAR, RBP_SAMPLING_EXTERNAL, RBP_FOLD_ORDERED, 16, 0, NULL, 0, &s);
    if (rbp_device_count() < 1) {
        printf("no device: status %d (%s)\n", status, rbp_status_string(status));
        return status == RBP_ERR_NO_DEVICE && s == NULL ? 0 : 2;
    }
    if (status != RBP_OK) return 3;
    status = rbp_solver_step(s, 64);
    rbp_solver_destroy(s);
    printf("device: status %d\n", status);
    return status == RBP_OK ? 0 : 4;
}
