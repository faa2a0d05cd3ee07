/* Compiled by tests/test_host.py with a plain C compiler: include/shb200.h must be valid C and libshb200.so must link
 * and run from C without torch/python.  Host-only calls here; the plan call is expected to fail cleanly with
 * SHB200_ECUDA on a machine without a GPU (no CPU fallback) or to succeed and run one tiny transform on a GPU box. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "shb200.h"

int main(void) {
    double x[8], w[8], sum = 0.0;
    if (shb200_version() != SHB200_VERSION) return 10;
    if (shb200_gauss_nodes(8, x, w) != SHB200_OK) return 11;
    for (int i = 0; i < 8; ++i) sum += w[i];
    if (fabs(sum - 2.0) > 1e-14) return 12;
    if (shb200_gauss_nodes(0, x, w) != SHB200_EINVAL) return 13;
    if (strlen(shb200_last_error()) == 0) return 14;

    double lat[4] = {-60.0, -20.0, 20.0, 60.0}, lon[8];
    for (int j = 0; j < 8; ++j) lon[j] = 45.0 * j;
    shb200_plan_desc d;
    memset(&d, 0, sizeof d);
    d.struct_size = (int32_t)sizeof d;
    d.nmax = 2; d.max_batch = 1;
    d.mode = SHB200_MODE_GRID; d.nlat = 4; d.lat_deg = lat; d.nlon = 8; d.lon_deg = lon;
    shb200_plan* plan = NULL;
    int rc = shb200_plan_create(&d, &plan);
    if (rc == SHB200_ECUDA) { printf("no-gpu: %s\n", shb200_last_error()); return 0; }
    if (rc != SHB200_OK) return 20;
    /* Y00 = 1 everywhere: coefficient vector (1,0,...,0) must synthesise to a grid of ones */
    double coef[9] = {1, 0, 0, 0, 0, 0, 0, 0, 0}, grid[32];
    rc = shb200_synthesis_host(plan, 1, coef, 9, 1, grid, 32, 8, 1, NULL);
    if (rc != SHB200_OK) { shb200_plan_destroy(plan); return 21; }
    for (int i = 0; i < 32; ++i)
        if (fabs(grid[i] - 1.0) > 1e-14) { shb200_plan_destroy(plan); return 22; }
    /* the same with a per-degree scale of 2 passed as a per-call option, into page-locked memory from the library's pool */
    double scale[3] = {2.0, 2.0, 2.0};
    shb200_call_opts o;
    memset(&o, 0, sizeof o);
    o.struct_size = (int32_t)sizeof o;
    o.degree_scale = scale;
    void* pin = NULL;
    if (shb200_host_alloc((int64_t)sizeof grid, &pin) != SHB200_OK) { shb200_plan_destroy(plan); return 23; }
    rc = shb200_synthesis_host(plan, 1, coef, 9, 1, (double*)pin, 32, 8, 1, &o);
    char names[128];
    shb200_plan_last_kernels(plan, names, (int32_t)sizeof names);
    shb200_plan_destroy(plan);
    if (rc != SHB200_OK) return 24;
    for (int i = 0; i < 32; ++i)
        if (fabs(((double*)pin)[i] - 2.0) > 1e-14) return 25;
    if (shb200_host_free(pin) != SHB200_OK || shb200_host_trim() != SHB200_OK) return 26;
    if (strstr(names, "k_pack") == NULL) return 27;
    printf("gpu: ok\n");
    return 0;
}
