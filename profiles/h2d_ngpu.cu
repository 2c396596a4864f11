// h2d_ngpu.cu -- bare pinned host-to-device copy bandwidth with 1, 2, 4, 8 GPUs copying CONCURRENTLY
// (is the 22.7 GB/s per GPU of the 8-GPU e2e leg the box or the chunk protocol of
// kfpos_batch_replay_toa?).  One host thread per GPU, one cudaMemcpyAsync per chunk, nothing else.
//
//   nvcc -O2 -o /tmp/h2d_ngpu profiles/h2d_ngpu.cu -lpthread && /tmp/h2d_ngpu > profiles/r02_h2d_ngpu.txt
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static double now() {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

struct Job {
    int dev;
    size_t chunk, total;
    unsigned flags;
    int streams;
    pthread_barrier_t *bar;
    double gbs;
};

static void *worker(void *arg) {
    Job *j = (Job *)arg;
    CK(cudaSetDevice(j->dev));
    void *h, *d;
    CK(cudaHostAlloc(&h, j->total, j->flags));
    memset(h, 1, j->total);
    CK(cudaMalloc(&d, j->total));
    cudaStream_t s[4];
    for (int i = 0; i < j->streams; ++i) CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
    // warm-up
    CK(cudaMemcpyAsync(d, h, j->total, cudaMemcpyHostToDevice, s[0]));
    CK(cudaStreamSynchronize(s[0]));
    pthread_barrier_wait(j->bar);
    const int reps = 8;
    const double t0 = now();
    for (int r = 0; r < reps; ++r) {
        int k = 0;
        for (size_t off = 0; off < j->total; off += j->chunk, ++k) {
            const size_t n = j->total - off < j->chunk ? j->total - off : j->chunk;
            CK(cudaMemcpyAsync((char *)d + off, (char *)h + off, n, cudaMemcpyHostToDevice, s[k % j->streams]));
        }
    }
    for (int i = 0; i < j->streams; ++i) CK(cudaStreamSynchronize(s[i]));
    const double t1 = now();
    j->gbs = (double)j->total * reps / (t1 - t0) / 1e9;
    pthread_barrier_wait(j->bar);
    CK(cudaFree(d));
    CK(cudaFreeHost(h));
    return 0;
}

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    printf("# bare pinned H2D copy, GB/s per GPU with n GPUs copying concurrently; 2 GiB per GPU x 8 passes\n");
    printf("# devices visible: %d\n", ndev);
    printf("%-6s %-10s %-8s %-14s %s\n", "n_gpus", "chunk_MB", "streams", "host_alloc", "GB/s per GPU (min .. max), total");
    const size_t total = (size_t)2 << 30;
    struct Cfg { size_t chunk; int streams; unsigned flags; const char *name; } cfgs[] = {
        {(size_t)256 << 20, 1, cudaHostAllocDefault, "default"},
        {(size_t)2 << 30, 1, cudaHostAllocDefault, "default"},
        {(size_t)32 << 20, 1, cudaHostAllocDefault, "default"},
        {(size_t)256 << 20, 2, cudaHostAllocDefault, "default"},
        {(size_t)256 << 20, 1, cudaHostAllocWriteCombined, "write-combined"},
        {(size_t)256 << 20, 1, cudaHostAllocPortable, "portable"},
    };
    for (int n = 1; n <= ndev && n <= 8; n *= 2) {
        for (unsigned c = 0; c < sizeof(cfgs) / sizeof(cfgs[0]); ++c) {
            pthread_barrier_t bar;
            pthread_barrier_init(&bar, 0, n);
            pthread_t th[8];
            Job jobs[8];
            for (int i = 0; i < n; ++i) {
                jobs[i].dev = i;
                jobs[i].chunk = cfgs[c].chunk;
                jobs[i].total = total;
                jobs[i].flags = cfgs[c].flags;
                jobs[i].streams = cfgs[c].streams;
                jobs[i].bar = &bar;
                jobs[i].gbs = 0;
                pthread_create(&th[i], 0, worker, &jobs[i]);
            }
            double mn = 1e30, mx = 0, sum = 0;
            for (int i = 0; i < n; ++i) {
                pthread_join(th[i], 0);
                mn = jobs[i].gbs < mn ? jobs[i].gbs : mn;
                mx = jobs[i].gbs > mx ? jobs[i].gbs : mx;
                sum += jobs[i].gbs;
            }
            pthread_barrier_destroy(&bar);
            printf("%-6d %-10zu %-8d %-14s %.1f .. %.1f, %.1f\n", n, cfgs[c].chunk >> 20, cfgs[c].streams, cfgs[c].name, mn, mx, sum);
            fflush(stdout);
        }
    }
    return 0;
}
