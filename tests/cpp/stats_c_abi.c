/* Plain C (not C++) caller of the boundary: creates a T6 batch, replays a few epochs, and reduces the error
 * statistics with kfpos_stats_allreduce -- the job's one collective -- exactly as a maintainer's C code would
 * (include/kfpos_b200.h).  Single rank: comm = NULL; the multi-rank path with a real
 * ncclComm_t is exercised by tests/test_gpu_stats.py.  Reads "N T M" and the tensors from stdin as text,
 * prints the six outputs with 17 digits.  Built and checked by tests/test_gpu_cpp_classes.py. */
#include <stdio.h>
#include <stdlib.h>

#include <kfpos_b200.h>

static double *read_doubles(size_t n) {
    double *p = (double *)malloc(sizeof(double) * n);
    for (size_t i = 0; i < n; ++i)
        if (scanf("%lf", &p[i]) != 1) exit(3);
    return p;
}

int main(void) {
    long N, T, M;
    if (scanf("%ld %ld %ld", &N, &T, &M) != 3) return 3;
    double *anchors = read_doubles(3 * (size_t)M), *x0 = read_doubles(6 * (size_t)N);
    double *truth = read_doubles(3 * (size_t)N), *ranges = read_doubles((size_t)(T * M * N));
    double *dt = (double *)malloc(sizeof(double) * (size_t)T);
    for (long t = 0; t < T; ++t) dt[t] = 0.1;
    kfpos_config cfg;
    kfpos_config_default(&cfg);
    cfg.accel_noise = 0.5;
    kfpos_batch *b = NULL;
    int rc = kfpos_batch_create(&b, 0, KFPOS_MODEL_T6, N, &cfg);
    if (rc) { fprintf(stderr, "create: %s\n", kfpos_strerror(rc)); return 1; }
    if ((rc = kfpos_batch_set_anchors(b, (int)M, anchors))) return 1;
    if ((rc = kfpos_batch_set_state(b, x0, NULL, NULL))) return 1;
    if ((rc = kfpos_batch_set_truth(b, truth, NULL))) return 1; /* the replay then ends with the block partials */
    if ((rc = kfpos_batch_replay_toa(b, (int)T, dt, ranges, KFPOS_FMT_F64_M, 0.01, NULL, NULL, NULL, NULL))) return 1;
    if ((rc = kfpos_batch_error_stats(b, NULL, NULL, NULL))) return 1; /* final tree, result stays on the device */
    double out[6];
    if ((rc = kfpos_stats_allreduce(b, (struct ncclComm *)0, NULL, out, NULL))) {
        fprintf(stderr, "allreduce: %s\n", kfpos_strerror(rc));
        return 1;
    }
    printf("stats %.17g %.17g %.17g %.17g %.17g %.17g\n", out[0], out[1], out[2], out[3], out[4], out[5]);
    kfpos_batch_destroy(b);
    return 0;
}
