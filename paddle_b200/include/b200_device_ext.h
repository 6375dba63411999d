/* Custom-device plug-in ABI of paddle_b200 (C, stable layout).
 *
 * A plug-in is a shared library that exports
 *     void B200InitPlugin(B200DeviceInterface* iface);
 * and fills the table below.  `struct_size` / `abi_version` are set by the LOADER before the call (so that an older plug-in can see how
 * much of the table it may write); the plug-in sets `device_type` and the callbacks it implements.  Required: initialize,
 * get_device_count, device_malloc, device_free, memcpy_h2d, memcpy_d2h, synchronize_device.  Every callback returns 0 on success.
 *
 * Role parity: paddle/phi/backends/device_ext.h (C_DeviceInterface / InitPlugin) - the contract of the reference's CustomDevice back ends.
 */
#ifndef B200_DEVICE_EXT_H_
#define B200_DEVICE_EXT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_DEVICE_ABI_VERSION 1

typedef int32_t B200Status; /* 0 = ok */
typedef void* B200Stream;
typedef void* B200Event;

/* one tensor argument of launch_kernel: device pointer + logical description */
typedef struct B200TensorArg {
  void* data;
  const char* dtype; /* "float32", "int64", ... */
  int32_t ndim;
  const int64_t* shape;
} B200TensorArg;

typedef struct B200DeviceInterface {
  size_t struct_size;   /* set by the loader: sizeof(B200DeviceInterface) it was built with */
  int32_t abi_version;  /* set by the loader */
  const char* device_type; /* set by the plug-in, e.g. "custom_cpu" */

  B200Status (*initialize)(void);
  B200Status (*finalize)(void);
  B200Status (*get_device_count)(int32_t* count);
  B200Status (*set_device)(int32_t device);

  B200Status (*device_malloc)(int32_t device, void** ptr, size_t size);
  B200Status (*device_free)(int32_t device, void* ptr, size_t size);
  B200Status (*memcpy_h2d)(int32_t device, void* dst, const void* src, size_t size);
  B200Status (*memcpy_d2h)(int32_t device, void* dst, const void* src, size_t size);
  B200Status (*memcpy_d2d)(int32_t device, void* dst, const void* src, size_t size);
  B200Status (*memory_stats)(int32_t device, size_t* total, size_t* free_bytes);

  B200Status (*create_stream)(int32_t device, B200Stream* stream);
  B200Status (*destroy_stream)(int32_t device, B200Stream stream);
  B200Status (*synchronize_stream)(int32_t device, B200Stream stream);
  B200Status (*create_event)(int32_t device, B200Event* event);
  B200Status (*record_event)(int32_t device, B200Stream stream, B200Event event);
  B200Status (*synchronize_event)(int32_t device, B200Event event);
  B200Status (*destroy_event)(int32_t device, B200Event event);
  B200Status (*synchronize_device)(int32_t device);

  /* optional: run operator `op` on device tensors (inputs first, then outputs); return non-zero for "not implemented" and the
     framework falls back to computing on the host and copying the result over */
  B200Status (*launch_kernel)(int32_t device, B200Stream stream, const char* op, const B200TensorArg* args, int32_t num_inputs, int32_t num_outputs);
} B200DeviceInterface;

void B200InitPlugin(B200DeviceInterface* iface);

#ifdef __cplusplus
}
#endif
#endif /* B200_DEVICE_EXT_H_ */
