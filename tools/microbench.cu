// Pipe-throughput microbenchmarks for the FP64 march: which per-step operations are cheap on sm_100a?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/bin/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 4;

template <int OP>
__global__ void __launch_bounds__(256) kern(double *out, double a, double b, int n) {
	double v[ILP];
	int iv[ILP];
	for (int i = 0; i < ILP; ++i) { v[i] = a + threadIdx.x * 1e-9 + i; iv[i] = threadIdx.x + i; }
	for (int it = 0; it < n; ++it) {
#pragma unroll
		for (int i = 0; i < ILP; ++i) {
			if (OP == 0) v[i] = __dadd_rn(v[i], b);
			if (OP == 1) v[i] = __dmul_rn(v[i], b);
			if (OP == 2) v[i] = __fma_rn(v[i], b, a);
			if (OP == 3) { iv[i] += __double2int_rz(v[i]); v[i] = __dadd_rn(v[i], b); }          // F2I + DADD
			if (OP == 4) { v[i] = __dadd_rn(v[i], (double)iv[i]); iv[i] += it; }                  // I2F + DADD
			if (OP == 5) { iv[i] += (v[i] < b) ? 1 : 2; v[i] = __dadd_rn(v[i], b); }              // DSETP + DADD
			if (OP == 6) v[i] = __ddiv_rn(v[i], b);
			if (OP == 7) { float f = (float)v[i]; iv[i] += __float_as_int(f); v[i] = __dadd_rn(v[i], b); } // F2F + DADD
			if (OP == 8) { iv[i] = iv[i] * 3 + it; }                                              // IMAD baseline
			if (OP == 9) { long long l = __double2ll_rz(v[i]); iv[i] += (int)l ^ (int)(l >> 32); v[i] = __dadd_rn(v[i], b); }
			if (OP == 10) { v[i] = __dsqrt_rn(v[i]); }
		}
	}
	double s = 0; int is = 0;
	for (int i = 0; i < ILP; ++i) { s += v[i]; is += iv[i]; }
	if (s == 12345.678 && is == 42) out[0] = s;
}

template <int OP>
void run(const char *name, int ops_per_iter) {
	double *d; CHECK(cudaMalloc(&d, 8));
	int dev; CHECK(cudaGetDevice(&dev));
	cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, dev));
	int blocks = p.multiProcessorCount * 8;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	kern<OP><<<blocks, 256>>>(d, 1.0000001, 0.99999, 16);
	CHECK(cudaDeviceSynchronize());
	cudaEventRecord(e0);
	kern<OP><<<blocks, 256>>>(d, 1.0000001, 0.99999, ITERS);
	cudaEventRecord(e1);
	CHECK(cudaDeviceSynchronize());
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double thread_iters = (double)blocks * 256 * ITERS * ILP;
	int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
	double per_sm_clk = thread_iters / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3);
	printf("%-28s %8.3f ms  %7.2f lane-iters/clk/SM (at %d MHz nominal)  -> %.2f Titer/s\n", name, ms, per_sm_clk, clk / 1000,
	       thread_iters / (ms * 1e-3) / 1e12);
	cudaFree(d);
}

int main() {
	cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
	printf("device %s sm_%d%d SMs %d L2 %d MB smem/SM %zu KB regs/SM %d maxThreads/SM %d\n", p.name, p.major, p.minor,
	       p.multiProcessorCount, p.l2CacheSize >> 20, p.sharedMemPerMultiprocessor >> 10, p.regsPerMultiprocessor,
	       p.maxThreadsPerMultiProcessor);
	run<0>("DADD", 1);
	run<1>("DMUL", 1);
	run<2>("DFMA", 1);
	run<3>("F2I.F64 + DADD", 2);
	run<4>("I2F.F64 + DADD", 2);
	run<5>("DSETP + DADD", 2);
	run<6>("DDIV", 1);
	run<7>("F2F.F32.F64 + DADD", 2);
	run<8>("IMAD", 1);
	run<9>("F2I.S64.F64 + DADD", 2);
	run<10>("DSQRT", 1);
	return 0;
}
