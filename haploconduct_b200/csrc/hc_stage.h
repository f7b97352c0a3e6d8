// hc_stage.h -- host<->device copies of large PAGEABLE buffers and stream-ordered device scratch for the
// host-buffer entry points (hc_fno1/3, hc_dedup_edges, hc_store_create*).  cudaMemcpy on pageable memory moves
// ~10 GB/s (one driver thread staging through one pinned buffer); the callers of the reference-facing ABI hold
// std::vector / numpy memory, so the library stages itself: several host threads copy into a ring of pinned
// buffers while the copy engine drains the previous one.  Pinned or registered buffers go straight to cudaMemcpy.
#ifndef HC_STAGE_H_
#define HC_STAGE_H_
#include <cuda_runtime.h>
#include <stddef.h>

// Synchronous like cudaMemcpy: ordered after all earlier work of the device, complete on return.
cudaError_t hc_copy_h2d(void* dst_dev, const void* src_host, size_t bytes);
cudaError_t hc_copy_d2h(void* dst_host, const void* src_dev, size_t bytes);

// The same through the ring, but with the device side of the copy on a stream of the caller's (the pipelines of
// hc_ingest_overlaps and hc_score_batch*): pinned / registered host memory is handed to cudaMemcpyAsync as it is.
//   hc_copy_h2d_on returns when the last piece has been handed to the copy engine: `src` may be reused, the copies
//                  complete in stream order on `st` (the caller orders its kernels behind them with an event);
//   hc_copy_d2h_on returns when `dst` holds the data if `dst` is pageable; for pinned memory it only enqueues
//                  (cudaMemcpyAsync semantics) -- *completed says which.
cudaError_t hc_copy_h2d_on(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st);
cudaError_t hc_copy_d2h_on(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st, bool* completed);

// A file straight to the device: several host threads pread pieces of it into the pinned ring, the copy engine drains
// them -- the host never holds more than the ring (32 MB) of the file.  Reads `bytes` from offset 0 of `fd`; *got = bytes
// that arrived (less if the file is shorter).  Synchronous: the data is on the device on return.
cudaError_t hc_copy_file_h2d(void* dst_dev, int fd, size_t bytes, size_t* got);

// Device scratch from the device's default memory pool, ordered on the legacy default stream (the stream the
// host-buffer entry points launch on); freed memory stays in the pool (up to 16 GB), so that a call does not pay
// cudaMalloc / cudaFree for each of its dozen temporaries.
cudaError_t hc_scratch_alloc(void** p, size_t bytes);
void hc_scratch_free(void* p);
// the same on a stream of the caller's (the store's streams are non-blocking: they do not order with the default stream)
cudaError_t hc_scratch_alloc_on(void** p, size_t bytes, cudaStream_t st);
void hc_scratch_free_on(void* p, cudaStream_t st);

#endif
