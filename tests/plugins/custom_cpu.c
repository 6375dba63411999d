/* Sample custom-device plug-in used by tests/test_custom_device_cpu.py: two "devices" backed by host memory with a byte budget, working
 * streams / events (counters), and ONE kernel ("add" on float32) so that both the device-kernel path and the host fallback are exercised. */
#include <stdlib.h>
#include <string.h>

#include "b200_device_ext.h"

#define N_DEV 2
#define BUDGET (64u << 20)
static size_t used[N_DEV];
static int initialized = 0, stream_count = 0, event_records = 0;

static B200Status init(void) { initialized = 1; return 0; }
static B200Status fini(void) { initialized = 0; return 0; }
static B200Status count(int32_t* n) { *n = N_DEV; return 0; }
static B200Status setdev(int32_t d) { return d >= 0 && d < N_DEV ? 0 : 2; }
static B200Status dmalloc(int32_t d, void** p, size_t size) {
  if (d < 0 || d >= N_DEV) return 2;
  if (used[d] + size > BUDGET) return 3; /* out of device memory */
  *p = malloc(size);
  if (!*p) return 3;
  used[d] += size;
  return 0;
}
static B200Status dfree(int32_t d, void* p, size_t size) { free(p); used[d] -= size; return 0; }
static B200Status cpy(int32_t d, void* dst, const void* src, size_t n) { (void)d; memcpy(dst, src, n); return 0; }
static B200Status mstats(int32_t d, size_t* total, size_t* fr) { *total = BUDGET; *fr = BUDGET - used[d]; return 0; }
static B200Status mkstream(int32_t d, B200Stream* s) { (void)d; *s = malloc(8); ++stream_count; return 0; }
static B200Status rmstream(int32_t d, B200Stream s) { (void)d; free(s); --stream_count; return 0; }
static B200Status syncstream(int32_t d, B200Stream s) { (void)d; (void)s; return 0; }
static B200Status mkevent(int32_t d, B200Event* e) { (void)d; *e = calloc(1, sizeof(int)); return 0; }
static B200Status recevent(int32_t d, B200Stream s, B200Event e) { (void)d; (void)s; *(int*)e = ++event_records; return 0; }
static B200Status syncevent(int32_t d, B200Event e) { (void)d; return *(int*)e > 0 ? 0 : 4; /* waiting on an event that was never recorded */ }
static B200Status rmevent(int32_t d, B200Event e) { (void)d; free(e); return 0; }
static B200Status syncdev(int32_t d) { (void)d; return initialized ? 0 : 5; }

static B200Status launch(int32_t d, B200Stream s, const char* op, const B200TensorArg* a, int32_t nin, int32_t nout) {
  (void)d; (void)s;
  if (strcmp(op, "add") == 0 && nin == 2 && nout == 1 && strcmp(a[0].dtype, "float32") == 0 && strcmp(a[1].dtype, "float32") == 0) {
    int64_t n = 1, n1 = 1;
    for (int i = 0; i < a[0].ndim; ++i) n *= a[0].shape[i];
    for (int i = 0; i < a[1].ndim; ++i) n1 *= a[1].shape[i];
    if (n != n1) return 1;
    const float* x = (const float*)a[0].data;
    const float* y = (const float*)a[1].data;
    float* o = (float*)a[2].data;
    for (int64_t i = 0; i < n; ++i) o[i] = x[i] + y[i];
    return 0;
  }
  return 1; /* not implemented: the framework falls back to the host */
}

void B200InitPlugin(B200DeviceInterface* f) {
  if (f->struct_size < sizeof(B200DeviceInterface)) { f->abi_version = -1; return; }
#ifdef PLUGIN_BAD_ABI
  f->abi_version = 99;
#endif
  f->device_type = "custom_cpu";
  f->initialize = init; f->finalize = fini; f->get_device_count = count; f->set_device = setdev;
  f->device_malloc = dmalloc; f->device_free = dfree; f->memcpy_h2d = cpy; f->memcpy_d2h = cpy; f->memcpy_d2d = cpy; f->memory_stats = mstats;
  f->create_stream = mkstream; f->destroy_stream = rmstream; f->synchronize_stream = syncstream;
  f->create_event = mkevent; f->record_event = recevent; f->synchronize_event = syncevent; f->destroy_event = rmevent;
  f->synchronize_device = syncdev;
#ifndef PLUGIN_NO_KERNELS
  f->launch_kernel = launch;
#endif
#ifdef PLUGIN_MISSING_REQUIRED
  f->memcpy_d2h = 0;
#endif
}
