// launch_floor.cu -- what a one-launch, flag-polled round trip costs on this box, as a function of the launch shape.
// Dev probe (not part of the library):  nvcc -arch=sm_100a -O3 -o launch_floor launch_floor.cu && ./launch_floor
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>

struct Small { float v[16]; };
struct Mid { float v[896]; };
struct Big { float v[2048]; };

template <typename P>
__global__ void __launch_bounds__(640, 1) probe_kernel(const __grid_constant__ P p, unsigned int* flag, unsigned int value, unsigned int* ticket,
                                                       float* sink, unsigned long long* stamps) {
    extern __shared__ float sm[];
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    if (threadIdx.x < 16) sm[threadIdx.x] = p.v[threadIdx.x];
    __syncthreads();
    __shared__ unsigned int last;
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) {
        *ticket = 0;
        sink[0] = sm[3];
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        stamps[0] = t0;
        stamps[1] = t1;
    }
}

template <typename P>
double run(const char* name, int grid, int threads, size_t smem, bool poll, int reps = 4000) {
    static P p;
    unsigned int *flag, *ticket;
    float* sink;
    unsigned long long* stamps;
    cudaMallocHost(&flag, 64);
    cudaMallocHost(&sink, 64);
    cudaMalloc(&ticket, 4);
    cudaMalloc(&stamps, 16);
    cudaMemset(ticket, 0, 4);
    cudaFuncSetAttribute(probe_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    double total = 0;
    for (int i = -200; i < reps; i++) {
        *(volatile unsigned int*)flag = 0;
        auto a = std::chrono::steady_clock::now();
        probe_kernel<P><<<grid, threads, smem, st>>>(p, flag, (unsigned)(i + 1000), ticket, sink, stamps);
        if (poll) {
            while (*(volatile unsigned int*)flag != (unsigned)(i + 1000)) {
            }
        } else {
            cudaStreamSynchronize(st);
        }
        auto b = std::chrono::steady_clock::now();
        if (i >= 0) total += std::chrono::duration<double, std::micro>(b - a).count();
    }
    cudaStreamSynchronize(st);
    unsigned long long hs[2];
    cudaMemcpy(hs, stamps, 16, cudaMemcpyDeviceToHost);
    printf("{\"probe\": \"%s\", \"grid\": %d, \"threads\": %d, \"smem\": %zu, \"param_bytes\": %zu, \"wait\": \"%s\", \"us_per_round_trip\": %.2f, \"kernel_first_cta_to_flag_us\": %.2f}\n",
           name, grid, threads, smem, sizeof(P), poll ? "poll" : "sync", total / reps, (hs[1] - hs[0]) / 1e3);
    return total / reps;
}

// SM clock actually seen by a short kernel launched in a tight host loop: clock64 ticks per globaltimer ns
__global__ void clock_probe(unsigned long long* out) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    while (clock64() - c0 < 20000) {
    }
    const long long c1 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    out[0] = (unsigned long long)(c1 - c0);
    out[1] = t1 - t0;
}

int main() {
    {
        unsigned long long* d;
        cudaMalloc(&d, 16);
        unsigned long long h[2] = {0, 0};
        double mhz_first = 0, mhz_last = 0;
        for (int i = 0; i < 2000; i++) {
            clock_probe<<<1, 1>>>(d);
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const double mhz = h[1] ? (double)h[0] / (double)h[1] * 1e3 : 0;
            if (i == 0) mhz_first = mhz;
            mhz_last = mhz;
        }
        printf("{\"probe\": \"SM clock seen by short kernels (clock64 / globaltimer)\", \"first_launch_mhz\": %.0f, \"after_2000_launches_mhz\": %.0f}\n", mhz_first, mhz_last);
    }
    run<Small>("1 CTA, small params", 1, 32, 0, true);
    run<Small>("1 CTA, small params", 1, 32, 0, false);
    run<Small>("148 CTAs x 640 thr, no smem", 148, 640, 0, true);
    run<Small>("148 CTAs x 640 thr, 200 KB smem", 148, 640, 200 * 1024, true);
    run<Mid>("148 CTAs x 640 thr, 200 KB smem, 3.5 KB params", 148, 640, 200 * 1024, true);
    run<Big>("148 CTAs x 640 thr, 200 KB smem, 8 KB params", 148, 640, 200 * 1024, true);
    run<Big>("148 CTAs x 640 thr, 200 KB smem, 8 KB params", 148, 640, 200 * 1024, false);
    run<Small>("148 CTAs x 256 thr, 64 KB smem", 148, 256, 64 * 1024, true);
    run<Small>("296 CTAs x 256 thr, 64 KB smem", 296, 256, 64 * 1024, true);
    return 0;
}
