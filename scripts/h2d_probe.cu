/*
 * h2d_probe.cu — what the platform gives host-to-device copies when several GPUs copy at once (the ceiling the end-to-end extraction
 * numbers of bench.py are quoted against), for several kinds of pinned host memory and several subsets of the GPUs.
 * One host thread per GPU, every thread copies its own host buffer to its own device buffer; all start together; reported per GPU
 * (CUDA events) and in aggregate (bytes of all GPUs / the slowest GPU's time).
 *   nvcc -O2 -std=c++17 -o scripts/_bin/h2d_probe scripts/h2d_probe.cu -lpthread
 *   scripts/_bin/h2d_probe [MiB per GPU, default 2048]
 */
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <sys/mman.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

enum Mode { DEFAULT = 0, WRITE_COMBINED = 1, HUGE_REGISTERED = 2, TWO_STREAMS = 3 };
static const char* mode_name[] = {"cudaHostAlloc (default)", "cudaHostAlloc write-combined", "2 MiB transparent huge pages + cudaHostRegister", "cudaHostAlloc, copy split over two streams"};

static std::atomic<int> arrived{0};
static void rendezvous(int n) { arrived.fetch_add(1); while (arrived.load() < n) {} }

static void worker(int dev, size_t bytes, Mode mode, int n_threads, int reps, double* gbs) {
    CK(cudaSetDevice(dev));
    void* d = nullptr; CK(cudaMalloc(&d, bytes));
    void* h = nullptr;
    if (mode == HUGE_REGISTERED) {
        h = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (h == MAP_FAILED) { printf("mmap failed\n"); exit(1); }
        madvise(h, bytes, MADV_HUGEPAGE);
        memset(h, 1, bytes);
        CK(cudaHostRegister(h, bytes, cudaHostRegisterDefault));
    } else {
        CK(cudaHostAlloc(&h, bytes, mode == WRITE_COMBINED ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
        memset(h, 1, bytes);
    }
    cudaStream_t s[2]; CK(cudaStreamCreate(&s[0])); CK(cudaStreamCreate(&s[1]));
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s[0])); CK(cudaDeviceSynchronize());      /* warm-up */
    rendezvous(n_threads);
    CK(cudaEventRecord(e0, s[0]));
    if (mode == TWO_STREAMS) CK(cudaStreamWaitEvent(s[1], e0, 0));
    for (int r = 0; r < reps; r++) {
        if (mode == TWO_STREAMS) {
            CK(cudaMemcpyAsync(d, h, bytes / 2, cudaMemcpyHostToDevice, s[0]));
            CK(cudaMemcpyAsync((char*)d + bytes / 2, (char*)h + bytes / 2, bytes - bytes / 2, cudaMemcpyHostToDevice, s[1]));
        } else CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s[0]));
    }
    if (mode == TWO_STREAMS) { CK(cudaEventRecord(e2, s[1])); CK(cudaStreamWaitEvent(s[0], e2, 0)); }
    CK(cudaEventRecord(e1, s[0]));
    CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    *gbs = (double)bytes * reps / (ms * 1e-3) / 1e9;
    if (mode == HUGE_REGISTERED) { CK(cudaHostUnregister(h)); munmap(h, bytes); } else CK(cudaFreeHost(h));
    CK(cudaFree(d));
}

static void run(const std::vector<int>& devs, size_t bytes, Mode mode) {
    arrived = 0;
    std::vector<double> gbs(devs.size(), 0.0);
    std::vector<std::thread> th;
    for (size_t i = 0; i < devs.size(); i++) th.emplace_back(worker, devs[i], bytes, mode, (int)devs.size(), 3, &gbs[i]);
    for (auto& t : th) t.join();
    double mn = 1e30, sum = 0; std::string list, each;
    for (size_t i = 0; i < devs.size(); i++) { mn = gbs[i] < mn ? gbs[i] : mn; sum += gbs[i]; list += std::to_string(devs[i]); char b[32]; snprintf(b, sizeof b, " %.1f", gbs[i]); each += b; }
    printf("GPUs %-8s %-50s per GPU%s GB/s, aggregate at the slowest GPU's pace %.1f GB/s (sum %.1f)\n", list.c_str(), mode_name[mode], each.c_str(), mn * devs.size(), sum);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 2048) << 20;
    int n = 0; CK(cudaGetDeviceCount(&n));
    printf("%d GPU(s), %zu MiB per GPU and copy, 3 copies per measurement\n", n, bytes >> 20);
    std::vector<std::vector<int>> sets;
    sets.push_back({0});
    if (n >= 2) sets.push_back({0, 1});
    if (n >= 4) { sets.push_back({0, 1, 2, 3}); }
    if (n >= 8) { sets.push_back({4, 5, 6, 7}); sets.push_back({0, 2, 4, 6}); sets.push_back({0, 1, 4, 5}); sets.push_back({0, 1, 2, 3, 4, 5, 6, 7}); }
    for (auto& s : sets) run(s, bytes, DEFAULT);
    std::vector<int> all; for (int i = 0; i < n; i++) all.push_back(i);
    for (Mode m : {WRITE_COMBINED, HUGE_REGISTERED, TWO_STREAMS}) { run({0}, bytes, m); if (n > 1) run(all, bytes, m); }
    return 0;
}
