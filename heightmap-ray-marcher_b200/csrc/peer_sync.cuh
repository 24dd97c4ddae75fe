// Device-side completion of a frame that several GPUs fill (include/hmrm.h "peer frames").
//
// The frame and a small control block live in the ROOT rank's device memory; every other rank maps both through
// CUDA IPC and its render kernel stores its tile rows straight into the root's frame over NVLink.  Round 1 told the
// root "the frame is complete" with a one-element NCCL all-reduce driven from Python (0.10 ms of a 0.34 ms frame at
// N = 8).  Here the ranks talk through two words of the control block instead, with tiny stream-ordered kernels:
//
//   every rank, use u of a buffer:   k_peer_spin(released >= u - 1)  ->  render kernel  ->  k_peer_signal(arrived += 1)
//   root:                            ... -> k_peer_spin(arrived >= u * ranks) -> [reads the frame] -> k_peer_release(u)
//
// (These kernels are the HMRM_PEER_SYNC=kernels implementation.  The default one does the same waits and writes on
// the same block with cuStreamWaitValue32 / cuStreamWriteValue32 — the GPU's front end, no SM; see hmrm_api.cu.)
// `arrived` only grows (one increment per rank and use), `released` is the last use the root is done with: a rank may
// overwrite the buffer for use u only when use u - 1 has been read.  No host synchronisation, no collective.
// A wait that does not end within the timeout raises the block's error word and lets the stream continue, so that a
// missing peer shows up as an error code instead of a hung GPU.
#ifndef HMRM_PEER_SYNC_CUH
#define HMRM_PEER_SYNC_CUH

#include <stdint.h>

namespace hmrm {

struct PeerCtrl {                 // HMRM_PEER_CTRL_BYTES, zero-initialised by hmrm_device_alloc
	unsigned int arrived;         // kernel protocol: one increment per rank and use
	unsigned int arrived_by[31];  // stream-memory-operation protocol: rank r writes its last finished use to [r]
	unsigned int released;        // on its own 128-byte line: written by the root, polled by the peers
	unsigned int error;
	unsigned int pad1[30];
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
	unsigned int v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// one thread: wait until *word has reached `target` (wrap-safe), or give up after timeout_ns
__global__ void k_peer_spin(const unsigned int *word, unsigned int target, unsigned int *error, unsigned long long timeout_ns) {
	const unsigned long long t0 = global_timer_ns();
	unsigned int backoff = 32u;
	while ((int)(ld_acquire_sys(word) - target) < 0) {
		__nanosleep(backoff);
		if (backoff < 1024u) backoff *= 2u;
		if (global_timer_ns() - t0 > timeout_ns) {
			atomicExch_system(error, 1u);
			break;
		}
	}
}

// one thread, stream-ordered after the render kernel: its stores (peer memory included) are made visible system-wide
// before the increment is
__global__ void k_peer_signal(unsigned int *arrived) {
	__threadfence_system();
	atomicAdd_system(arrived, 1u);
}

__global__ void k_peer_release(unsigned int *released, unsigned int use) {
	__threadfence_system();
	atomicExch_system(released, use);
}

} // namespace hmrm

#endif
