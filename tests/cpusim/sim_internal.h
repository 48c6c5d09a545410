// tests/cpusim — interfaces between the simulator's translation units (TEST INFRASTRUCTURE ONLY, see sim_device.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <vector>

namespace cpusim {

// ---- stream scheduler (sim_cudart.cxx) ---------------------------------------------------------------------------------
// CPUSIM_SCHED=sync (default): work executes when it is enqueued.
// CPUSIM_SCHED=fifo | lifo | random:<seed>: work is only QUEUED on its stream and runs when the host reaches a
// synchronisation point (cudaStreamSynchronize, cudaDeviceSynchronize, cudaEventSynchronize, cudaFree, a copy that CUDA
// defines as host-synchronous); the scheduler then picks among the streams whose head task is runnable — in enqueue order
// (fifo), preferring the task enqueued LAST (lifo: everything that is not ordered by an event or by stream order runs in
// the "wrong" order) or at random.  A host schedule that forgot an event dependency between two streams computes garbage
// under lifo/random while it may pass on a GPU for years.
void stream_submit(cudaStream_t s, std::function<void()> fn);

// ---- simulated NCCL (sim_nccl.cxx) -------------------------------------------------------------------------------------
struct NcclBatch;  // the operations one ncclGroup (or one ungrouped call) issued on one stream
void stream_submit_nccl(cudaStream_t s, NcclBatch* b);
bool nccl_progress(std::vector<NcclBatch*>& active);  // one nonblocking pass over every active batch; true if anything moved
bool nccl_done(const NcclBatch* b);
void nccl_finish(NcclBatch* b);    // reductions + release
void nccl_describe(const NcclBatch* b);
void run_batch_blocking(NcclBatch* b);   // host-side setup exchanges (communicator creation)

// what the simulator's cuTensorMapEncodeTiled writes into the opaque 128-byte CUtensorMap (2-D, FP64 only)
struct SimTensorMap {
  uint64_t magic;
  const char* base;
  uint64_t dim[2];       // elements; dim[0] is contiguous
  uint64_t stride1;      // bytes between consecutive dim-1 indices
  uint32_t box[2];
  uint32_t swizzle128;
  uint32_t elem_bytes;   // 8 (FLOAT64) or 4 (FLOAT32)
};
constexpr uint64_t kTensorMapMagic = 0x53494d544d415031ull;

}  // namespace cpusim
