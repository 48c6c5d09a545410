// candmc_b200 — host<->device staging for the C ABI: the reference's callers own HOST matrices
// (posix_memalign in test/MM/topo_pdgemm_unit.cxx:226-248); device pointers take the zero-copy fast path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "common.cuh"

namespace candmc {

inline bool is_device_ptr(const void* p) {
  if (p == nullptr) return true;  // nothing to stage
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// A column-major matrix operand that lives on the device for the duration of one call.
// If the caller's pointer is a host pointer the data is packed (ld = rows) into a temporary device buffer.
class StagedMatrix {
 public:
  StagedMatrix() = default;
  StagedMatrix(const StagedMatrix&) = delete;
  StagedMatrix& operator=(const StagedMatrix&) = delete;
  ~StagedMatrix() {
    if (owned_) cudaFree(owned_);
  }

  // copy_in: host data is needed on the device (inputs, or C when beta != 0)
  int open(const double* user, int64_t rows, int64_t cols, int64_t ld, bool copy_in, cudaStream_t stream) {
    user_ = const_cast<double*>(user);
    rows_ = rows;
    cols_ = cols;
    user_ld_ = ld;
    if (user == nullptr || rows == 0 || cols == 0 || is_device_ptr(user)) {
      dev_ = user_;
      ld_ = ld;
      staged_ = false;
      return OK;
    }
    staged_ = true;
    ld_ = rows;
    cudaError_t e = cudaMalloc(&owned_, sizeof(double) * rows * cols);
    if (e != cudaSuccess) {
      set_last_error("staging: cudaMalloc(%lld x %lld doubles) failed: %s", (long long)rows, (long long)cols,
                     cudaGetErrorString(e));
      return ERR_NOMEM;
    }
    dev_ = static_cast<double*>(owned_);
    if (copy_in) {
      CANDMC_CUDA(cudaMemcpy2DAsync(dev_, ld_ * 8, user_, user_ld_ * 8, rows * 8, cols, cudaMemcpyHostToDevice,
                                    stream));
    }
    return OK;
  }
  // write the device copy back to the caller's host matrix (no-op for device operands)
  int close_out(cudaStream_t stream) {
    if (!staged_) return OK;
    CANDMC_CUDA(cudaMemcpy2DAsync(user_, user_ld_ * 8, dev_, ld_ * 8, rows_ * 8, cols_, cudaMemcpyDeviceToHost,
                                  stream));
    return OK;
  }
  // columns [c0, c0 + w) of a device matrix `src` (leading dimension ld_src) to the same columns of the caller's host matrix
  int close_out_cols(const double* src, int64_t ld_src, int64_t c0, int64_t w, cudaStream_t stream) {
    if (!staged_ || w <= 0) return OK;
    CANDMC_CUDA(cudaMemcpy2DAsync(user_ + c0 * user_ld_, user_ld_ * 8, src, ld_src * 8, rows_ * 8, w, cudaMemcpyDeviceToHost,
                                  stream));
    return OK;
  }
  double* ptr() const { return dev_; }
  int64_t ld() const { return ld_; }
  bool staged() const { return staged_; }

 private:
  double* user_ = nullptr;
  double* dev_ = nullptr;
  void* owned_ = nullptr;
  int64_t rows_ = 0, cols_ = 0, ld_ = 0, user_ld_ = 0;
  bool staged_ = false;
};

}  // namespace candmc
