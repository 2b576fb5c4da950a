// TEST INFRASTRUCTURE ONLY (tests/sanitize/): a stand-in for <cuda_runtime.h> so that the HOST code of
// libperseus-sdr_b200/csrc/{handle,stream_path,bulk_path}.cu -- handle lock, slab ring, latency watchdog, staging pipeline bookkeeping --
// can be compiled with a plain C++ compiler and run under ThreadSanitizer / AddressSanitizer (tests/sanitize/sanitize.sh).
// "Device" memory is host memory; copies and kernels complete before the call returns; host functions (cudaLaunchHostFunc) run
// on a thread of their own, in order, and events / stream synchronisation wait for the ones queued before them.
// Nothing here is part of the product, and the product never compiles against it.
#pragma once
#include <cstddef>
#include <cstdint>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
typedef struct fake_stream *cudaStream_t;
typedef struct fake_event *cudaEvent_t;
enum { cudaStreamNonBlocking = 1 };
enum { cudaEventDisableTiming = 2, cudaEventBlockingSync = 1 };
enum { cudaHostAllocDefault = 0, cudaHostAllocPortable = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; };
struct cudaDeviceProp { char name[256]; int multiProcessorCount, major, minor; size_t totalGlobalMem; };

cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDevice(int *);
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *, int);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
#define CUDART_CB
typedef void (*cudaHostFn_t)(void *);
cudaError_t cudaLaunchHostFunc(cudaStream_t, cudaHostFn_t, void *);
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *, unsigned);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventQuery(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float *, cudaEvent_t, cudaEvent_t);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaMalloc(void **, size_t);
cudaError_t cudaFree(void *);
cudaError_t cudaHostAlloc(void **, size_t, unsigned);
cudaError_t cudaFreeHost(void *);
cudaError_t cudaMemcpyAsync(void *, const void *, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t);
cudaError_t cudaMemset(void *, int, size_t);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *, const void *);

template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc(reinterpret_cast<void **>(p), n); }
template <class T> inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned f) { return cudaHostAlloc(reinterpret_cast<void **>(p), n, f); }
