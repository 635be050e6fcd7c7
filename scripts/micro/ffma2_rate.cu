// Microbenchmark: issue rate of FFMA (scalar) vs FFMA2 (packed f32x2) on sm_100a, alone and mixed with ALU-pipe work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CH 8

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int n_alu_dummy) {
    float2 v[CH];
    for (int c = 0; c < CH; c++) v[c] = make_float2(threadIdx.x*0.001f + c, threadIdx.x*0.002f - c);
    unsigned w[CH];
    for (int c = 0; c < CH; c++) w[c] = threadIdx.x*(c + 1) + n_alu_dummy;
    const float2 a2 = make_float2(a, a*1.0001f), b2 = make_float2(b, b*0.9999f);
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (MODE == 0 || MODE == 2) {          // scalar: two FFMA per float2
                v[c].x = fmaf(v[c].x, a2.x, b2.x);
                v[c].y = fmaf(v[c].y, a2.y, b2.y);
            } else if (MODE == 1 || MODE == 3) {    // packed: one FFMA2
                v[c] = __ffma2_rn(v[c], a2, b2);
            } else if (MODE == 4) {
                v[c] = __fmul2_rn(v[c], a2);
            } else if (MODE == 5) {
                v[c] = __fadd2_rn(v[c], b2);
            } else if (MODE == 6) {                 // scalar FMUL x2
                v[c].x = v[c].x*a2.x; v[c].y = v[c].y*a2.y;
            } else if (MODE == 7) {                 // FFMA2 with a broadcast scalar operand
                v[c] = __ffma2_rn(v[c], make_float2(a, a), b2);
            }
            if (MODE == 2 || MODE == 3) {           // plus two ALU-pipe instructions per float2
                w[c] = (w[c] ^ (w[c] >> 3)) + 0x9e3779b9u;     // LOP3/SHF + IADD
            }
        }
    }
    float s = 0.f; unsigned t = 0;
    for (int c = 0; c < CH; c++) { s += v[c].x + v[c].y; t ^= w[c]; }
    out[blockIdx.x*blockDim.x + threadIdx.x] = s + (float)t;
}

// dependent-issue latency: one warp, one chain
template <int MODE>
__global__ void lat(float* out, long long* cycles, float a, float b) {
    float2 v = make_float2(threadIdx.x*0.001f, threadIdx.x*0.002f);
    const float2 a2 = make_float2(a, a*1.0001f), b2 = make_float2(b, b*0.9999f);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; it++) {
#pragma unroll
        for (int c = 0; c < 16; c++) {
            if (MODE == 0) v.x = fmaf(v.x, a2.x, b2.x);
            else if (MODE == 1) v = __ffma2_rn(v, a2, b2);
            else if (MODE == 2) v = __fmul2_rn(v, a2);
            else v = __fadd2_rn(v, b2);
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = v.x + v.y;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE>
void run_lat(const char* name, float* out) {
    long long* c; cudaMalloc(&c, 8);
    lat<MODE><<<1, 32>>>(out, c, 1.0001f, 0.5f);
    lat<MODE><<<1, 32>>>(out, c, 1.0001f, 0.5f);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-28s dependent-issue latency %.2f cycles\n", name, (double)h/(256*16));
    cudaFree(c);
}

template <int MODE>
void run(const char* name, float* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148*8;
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0*2*CH*(double)ITERS*blocks*256;       // per float2: 2 fma = 4 flop
    printf("%-28s %8.3f ms  %7.2f TFLOP/s fp32  (%.1f G float2-fma/s)\n", name, ms, flops/ms*1e-9, flops/4/ms*1e-6);
}

int main() {
    float* out; cudaMalloc(&out, sizeof(float)*148*8*256);
    run<0>("scalar FFMA x2", out);
    run<1>("packed FFMA2", out);
    run<2>("scalar FFMA x2 + ALU", out);
    run<3>("packed FFMA2 + ALU", out);
    run<4>("packed FMUL2", out);
    run<5>("packed FADD2", out);
    run<6>("scalar FMUL x2", out);
    run<7>("FFMA2 broadcast operand", out);
    run_lat<0>("FFMA chain", out);
    run_lat<1>("FFMA2 chain", out);
    run_lat<2>("FMUL2 chain", out);
    run_lat<3>("FADD2 chain", out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
