// Prototype measurement (diagnostic, not product) behind DESIGN.md's decision on the K-sharded single fit (SURVEY 8(e),
// BASELINE.json configs[4]): the per-neuron-step all-reduce of (P + 1) doubles between G GPUs of one box, device
// initiated over NVLink peer memory, as the K-sharded sweep would need it once per chain step (N x iters sequential
// exchanges, caviar.py:196-229).
//
// One persistent CTA per GPU (single process, one host thread per device).  Every step each rank
//   1. computes a dummy partial (P + 1 doubles),
//   2. stores it into slot (step & 1) of every peer's mailbox with a release store of the step number as flag,
//   3. spins (ld.acquire.sys) until all G - 1 peer flags of that slot carry the step number, sums in rank order.
// Reports microseconds per exchange; compare with the measured chain-step time of a single-GPU fit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_exchange peer_exchange.cu -Xcompiler -pthread
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

constexpr int PV = 4;            // P + 1 values
constexpr int MAXG = 8;
struct Mail { double v[2][MAXG][PV]; int flag[2][MAXG]; int pad[16]; };

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) { int v; asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

struct Peers { Mail* box[MAXG]; };

__global__ void exchange_kernel(Peers peers, int rank, int G, int steps, double* out, long long* cycles) {
    const int lane = threadIdx.x;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int s = 1; s <= steps; ++s) {
        const int slot = s & 1;
        const double part = 1e-3 * (rank + 1) + 1e-6 * s + acc * 1e-12;       // depends on the previous result: sequential chain
        // lanes 0..PV-1 carry the values, lane r < G delivers to peer r
        if (lane < G && lane != rank) {
            Mail* dst = peers.box[lane];
            for (int q = 0; q < PV; ++q) dst->v[slot][rank][q] = part + q;
            st_release_sys(&dst->flag[slot][rank], s);
        }
        double tot = 0.0;
        if (lane == 0) {
            Mail* me = peers.box[rank];
            for (int r = 0; r < G; ++r) {
                if (r == rank) { tot += part; continue; }
                while (ld_acquire_sys(&me->flag[slot][r]) < s) {}
                tot += me->v[slot][r][0];
            }
        }
        tot = __shfl_sync(0xffffffffu, tot, 0);
        acc = tot;
    }
    if (lane == 0) { *cycles = clock64() - t0; *out = acc; }
}

int main(int argc, char** argv) {
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    int G = argc > 1 ? atoi(argv[1]) : ndev;
    if (G > ndev) G = ndev;
    if (G < 2) { printf("needs >= 2 GPUs (found %d)\n", ndev); return 0; }
    const int steps = argc > 2 ? atoi(argv[2]) : 56612;          // chain steps of one C5 fit
    Peers peers{};
    for (int d = 0; d < G; ++d) {
        cudaSetDevice(d);
        for (int e = 0; e < G; ++e) if (e != d) cudaDeviceEnablePeerAccess(e, 0);
        cudaMalloc(&peers.box[d], sizeof(Mail));
        cudaMemset(peers.box[d], 0, sizeof(Mail));
    }
    std::vector<double*> outs(G); std::vector<long long*> cyc(G);
    for (int d = 0; d < G; ++d) { cudaSetDevice(d); cudaMalloc(&outs[d], 8); cudaMalloc(&cyc[d], 8); cudaDeviceSynchronize(); }
    for (int rep = 0; rep < 2; ++rep) {
        for (int d = 0; d < G; ++d) { cudaSetDevice(d); cudaMemset(peers.box[d], 0, sizeof(Mail)); cudaDeviceSynchronize(); }
        std::vector<std::thread> th;
        std::vector<float> ms(G);
        for (int d = 0; d < G; ++d)
            th.emplace_back([&, d] {
                cudaSetDevice(d);
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                exchange_kernel<<<1, 32>>>(peers, d, G, steps, outs[d], cyc[d]);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms[d], e0, e1);
            });
        for (auto& t : th) t.join();
        float mx = 0; for (float m : ms) mx = m > mx ? m : mx;
        double o; cudaSetDevice(0); cudaMemcpy(&o, outs[0], 8, cudaMemcpyDeviceToHost);
        printf("G=%d ranks, %d sequential exchanges of %d doubles: %.2f ms total, %.3f us per exchange (result %.6f, err %s)\n", G, steps, PV,
               mx, 1e3 * mx / steps, o, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
