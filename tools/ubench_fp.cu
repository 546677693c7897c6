// Micro-benchmark: issue rate of unfused FP32 multiply+add, scalar vs packed (.f32x2), on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp ubench_fp.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

constexpr int ILP = 16;
template <int MODE> __global__ void k(float* out, const float* in, int iters) {
    float w0 = in[0], w1 = in[1], x0 = in[2 + threadIdx.x % 4], x1 = in[8 + threadIdx.x % 3];
    if (MODE == 0) {          // scalar FMUL + FADD
        float acc[ILP];
        float xs[4] = {x0, x1, x0 + 2.f, x1 + 3.f}, ws[4] = {w0, w1, in[20], in[21]};
        for (int i = 0; i < ILP; i++) acc[i] = in[i];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = __fadd_rn(acc[i], __fmul_rn(xs[i & 3], ws[i >> 2]));
#pragma unroll
            for (int j = 0; j < 4; j++) xs[j] = __fadd_rn(xs[j], 1.0f);
        }
        float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 1) {   // packed: ILP/2 pairs
        uint64_t acc[ILP / 2];
        for (int i = 0; i < ILP / 2; i++) acc[i] = pk(in[2 * i], in[2 * i + 1]);
        uint64_t xp[2] = {pk(x0, x1), pk(x1 + 2.f, x0 + 3.f)}, wp[4] = {pk(w0, w0), pk(w1, w1), pk(in[20], in[20]), pk(in[21], in[21])};
        const uint64_t one = pk(1.0f, 1.0f);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP / 2; i++) acc[i] = add2(acc[i], mul2(xp[i & 1], wp[i >> 1]));
            xp[0] = add2(xp[0], one); xp[1] = add2(xp[1], one);
        }
        float s = 0;
        for (int i = 0; i < ILP / 2; i++) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[i])); s += a + b; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 2) {   // scalar FFMA (reference peak)
        float acc[ILP];
        for (int i = 0; i < ILP; i++) acc[i] = in[i];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = __fmaf_rn((i & 1) ? x0 : x1, (i & 2) ? w0 : w1, acc[i]);
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = __fmaf_rn((i & 1) ? x1 : x0, (i & 2) ? w0 : w1, acc[i]);
            x0 += 1.0f; x1 += 1.0f;
        }
        float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 3) {   // FMNMX chain
        float acc[ILP];
        for (int i = 0; i < ILP; i++) acc[i] = in[i];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = fmaxf(acc[i], (i & 1) ? x0 : x1);
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = fminf(acc[i], (i & 2) ? x0 : x1);
            x0 += 1.0f; x1 += 1.0f;
        }
        float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 8) {   // 3-input FMNMX3 chain (counted as 2 ops each)
        float acc[ILP];
        for (int i = 0; i < ILP; i++) acc[i] = in[i];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) asm("max.NaN.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"((i & 1) ? x0 : x1), "f"((i & 2) ? x1 : x0));
            x0 += 1.0f; x1 += 1.0f;
        }
        float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 4) {   // scalar FADD only
        float acc[ILP];
        for (int i = 0; i < ILP; i++) acc[i] = in[i];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = __fadd_rn(acc[i], (i & 1) ? x0 : x1);
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = __fadd_rn(acc[i], (i & 2) ? x0 : x1);
            x0 += 1.0f; x1 += 1.0f;
        }
        float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 6) {   // unfused packed: fma2(x, w, runtime -0.0) then add2
        uint64_t acc[ILP / 2];
        for (int i = 0; i < ILP / 2; i++) acc[i] = pk(in[2 * i], in[2 * i + 1]);
        uint64_t xp[2] = {pk(x0, x1), pk(x1 + 2.f, x0 + 3.f)}, wp[4] = {pk(w0, w0), pk(w1, w1), pk(in[20], in[20]), pk(in[21], in[21])};
        const uint64_t one = pk(1.0f, 1.0f);
        const uint64_t nz = pk(in[30], in[30]);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP / 2; i++) acc[i] = add2(acc[i], fma2(xp[i & 1], wp[i >> 1], nz));
            xp[0] = add2(xp[0], one); xp[1] = add2(xp[1], one);
        }
        float s = 0;
        for (int i = 0; i < ILP / 2; i++) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[i])); s += a + b; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 7) {   // scalar FMUL feeding packed FADD2
        uint64_t acc[ILP / 2];
        float xs[4] = {x0, x1, x0 + 2.f, x1 + 3.f}, ws[4] = {w0, w1, in[20], in[21]};
        for (int i = 0; i < ILP / 2; i++) acc[i] = pk(in[2 * i], in[2 * i + 1]);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP / 2; i++) acc[i] = add2(acc[i], pk(__fmul_rn(xs[(2 * i) & 3], ws[i >> 1]), __fmul_rn(xs[(2 * i + 1) & 3], ws[i >> 1])));
#pragma unroll
            for (int j = 0; j < 4; j++) xs[j] = __fadd_rn(xs[j], 1.0f);
        }
        float s = 0;
        for (int i = 0; i < ILP / 2; i++) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[i])); s += a + b; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 5) {   // packed add only
        uint64_t acc[ILP / 2];
        for (int i = 0; i < ILP / 2; i++) acc[i] = pk(in[2 * i], in[2 * i + 1]);
        uint64_t xa = pk(x0, x1), xb = pk(x1, x0);
        const uint64_t one = pk(1.0f, 1.0f);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP / 2; i++) acc[i] = add2(acc[i], (i & 1) ? xa : xb);
#pragma unroll
            for (int i = 0; i < ILP / 2; i++) acc[i] = add2(acc[i], (i & 2) ? xa : xb);
            xa = add2(xa, one);
        }
        float s = 0;
        for (int i = 0; i < ILP / 2; i++) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[i])); s += a + b; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

template <int MODE> void run(const char* name, float* out, float* in, int warps) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    const int iters = 20000, blocks = pr.multiProcessorCount, threads = warps * 32;
    k<MODE><<<blocks, threads>>>(out, in, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // scalar flops per thread per iter: 32 (16 mul + 16 add, or 32 of the op)
    double ops = (double)blocks * threads * iters * 32.0;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("%-22s warps/SM=%2d  %.1f Gop/s  = %.1f ops/clk/SM at %d MHz (err=%s)\n", name, warps, ops / ms / 1e6,
           ops / (ms * 1e-3) / blocks / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *out, *in; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    for (int warps : {8, 16, 32}) {
        run<0>("FMUL+FADD scalar", out, in, warps);
        run<1>("mul.f32x2+add.f32x2", out, in, warps);
        run<2>("FFMA scalar", out, in, warps);
        run<3>("FMNMX", out, in, warps);
        run<4>("FADD scalar", out, in, warps);
        run<5>("add.f32x2", out, in, warps);
        run<6>("fma2(x,w,-0)+add2", out, in, warps);
        run<7>("FMUL + add2", out, in, warps);
        run<8>("FMNMX3 (x0.5: 16/iter)", out, in, warps);
    }
    return 0;
}
