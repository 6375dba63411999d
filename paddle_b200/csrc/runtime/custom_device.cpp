// Loader and runtime for custom-device plug-ins (include/b200_device_ext.h).
// Parity (role): paddle/phi/backends/custom/custom_device.cc + device_manager.cc (LoadCustomRuntimeLib: dlopen, InitPlugin, validate the
// interface, register the device type) and the CustomDevice wrappers for memory, streams, events.
#include <dlfcn.h>

#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200_device_ext.h"
#include "runtime.h"

namespace py = pybind11;

namespace b200 {
namespace runtime {

class CustomDevice {
 public:
  explicit CustomDevice(const std::string& path) : path_(path) {
    handle_ = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!handle_) throw std::runtime_error(std::string("custom device: cannot load '") + path + "': " + dlerror());
    using InitFn = void (*)(B200DeviceInterface*);
    auto init = (InitFn)dlsym(handle_, "B200InitPlugin");
    if (!init) { dlclose(handle_); throw std::runtime_error("custom device: '" + path + "' does not export B200InitPlugin"); }
    std::memset(&iface_, 0, sizeof(iface_));
    iface_.struct_size = sizeof(iface_);
    iface_.abi_version = B200_DEVICE_ABI_VERSION;
    init(&iface_);
    if (iface_.abi_version != B200_DEVICE_ABI_VERSION || iface_.struct_size != sizeof(iface_)) {
      dlclose(handle_);
      throw std::runtime_error("custom device: '" + path + "' was built against another ABI version of b200_device_ext.h");
    }
    if (!iface_.device_type || !*iface_.device_type) { dlclose(handle_); throw std::runtime_error("custom device: the plug-in did not set device_type"); }
    const char* missing = nullptr;
    if (!iface_.initialize) missing = "initialize";
    else if (!iface_.get_device_count) missing = "get_device_count";
    else if (!iface_.device_malloc) missing = "device_malloc";
    else if (!iface_.device_free) missing = "device_free";
    else if (!iface_.memcpy_h2d) missing = "memcpy_h2d";
    else if (!iface_.memcpy_d2h) missing = "memcpy_d2h";
    else if (!iface_.synchronize_device) missing = "synchronize_device";
    if (missing) { dlclose(handle_); throw std::runtime_error(std::string("custom device: required callback '") + missing + "' is missing"); }
    type_ = iface_.device_type;
    check(iface_.initialize(), "initialize");
  }
  ~CustomDevice() {
    if (handle_) {
      for (auto& kv : live_)
        iface_.device_free(kv.second.device, (void*)kv.first, kv.second.size);
      if (iface_.finalize) iface_.finalize();
      dlclose(handle_);
    }
  }
  CustomDevice(const CustomDevice&) = delete;

  const std::string& type() const { return type_; }
  const std::string& path() const { return path_; }
  int device_count() const { int32_t n = 0; check(iface_.get_device_count(&n), "get_device_count"); return n; }
  void set_device(int d) const { if (iface_.set_device) check(iface_.set_device(d), "set_device"); }

  uint64_t malloc(int dev, size_t size) {
    void* p = nullptr;
    check(iface_.device_malloc(dev, &p, size), "device_malloc");
    if (!p && size) throw std::runtime_error("custom device: device_malloc returned NULL");
    std::lock_guard<std::mutex> g(mu_);
    live_[(uintptr_t)p] = {dev, size};
    allocated_ += size;
    peak_ = std::max(peak_, allocated_);
    return (uint64_t)(uintptr_t)p;
  }
  void free(uint64_t ptr) {
    Alloc a;
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = live_.find((uintptr_t)ptr);
      if (it == live_.end()) throw std::runtime_error("custom device: free of an unknown pointer");
      a = it->second;
      live_.erase(it);
      allocated_ -= a.size;
    }
    check(iface_.device_free(a.device, (void*)(uintptr_t)ptr, a.size), "device_free");
  }
  void h2d(int dev, uint64_t dst, const py::buffer& src) {
    py::buffer_info info = src.request();
    const size_t n = (size_t)info.size * (size_t)info.itemsize;
    bounds(dst, n);
    check(iface_.memcpy_h2d(dev, (void*)(uintptr_t)dst, info.ptr, n), "memcpy_h2d");
  }
  py::bytes d2h(int dev, uint64_t src, size_t n) {
    bounds(src, n);
    std::string out(n, '\0');
    check(iface_.memcpy_d2h(dev, out.data(), (const void*)(uintptr_t)src, n), "memcpy_d2h");
    return py::bytes(out);
  }
  void d2d(int dev, uint64_t dst, uint64_t src, size_t n) {
    bounds(dst, n);
    bounds(src, n);
    if (!iface_.memcpy_d2d) throw std::runtime_error("custom device: the plug-in has no memcpy_d2d");
    check(iface_.memcpy_d2d(dev, (void*)(uintptr_t)dst, (const void*)(uintptr_t)src, n), "memcpy_d2d");
  }
  py::tuple memory_stats(int dev) {
    size_t total = 0, fr = 0;
    if (iface_.memory_stats) check(iface_.memory_stats(dev, &total, &fr), "memory_stats");
    std::lock_guard<std::mutex> g(mu_);
    return py::make_tuple(total, fr, allocated_, peak_);
  }
  uint64_t create_stream(int dev) { B200Stream s = nullptr; need(iface_.create_stream, "create_stream"); check(iface_.create_stream(dev, &s), "create_stream"); return (uint64_t)(uintptr_t)s; }
  void destroy_stream(int dev, uint64_t s) { need(iface_.destroy_stream, "destroy_stream"); check(iface_.destroy_stream(dev, (B200Stream)(uintptr_t)s), "destroy_stream"); }
  void sync_stream(int dev, uint64_t s) { need(iface_.synchronize_stream, "synchronize_stream"); check(iface_.synchronize_stream(dev, (B200Stream)(uintptr_t)s), "synchronize_stream"); }
  uint64_t create_event(int dev) { B200Event e = nullptr; need(iface_.create_event, "create_event"); check(iface_.create_event(dev, &e), "create_event"); return (uint64_t)(uintptr_t)e; }
  void record_event(int dev, uint64_t s, uint64_t e) { need(iface_.record_event, "record_event"); check(iface_.record_event(dev, (B200Stream)(uintptr_t)s, (B200Event)(uintptr_t)e), "record_event"); }
  void sync_event(int dev, uint64_t e) { need(iface_.synchronize_event, "synchronize_event"); check(iface_.synchronize_event(dev, (B200Event)(uintptr_t)e), "synchronize_event"); }
  void destroy_event(int dev, uint64_t e) { need(iface_.destroy_event, "destroy_event"); check(iface_.destroy_event(dev, (B200Event)(uintptr_t)e), "destroy_event"); }
  void synchronize(int dev) { check(iface_.synchronize_device(dev), "synchronize_device"); }

  // args: [(ptr, dtype, shape)], inputs first.  Returns False when the plug-in has no kernel for `op` (caller falls back to the host).
  bool launch(int dev, uint64_t stream, const std::string& op, const py::list& args, int num_inputs) {
    if (!iface_.launch_kernel) return false;
    std::vector<B200TensorArg> a(args.size());
    std::vector<std::string> dts(args.size());
    std::vector<std::vector<int64_t>> shapes(args.size());
    for (size_t i = 0; i < args.size(); ++i) {
      auto t = args[i].cast<py::tuple>();
      a[i].data = (void*)(uintptr_t)t[0].cast<uint64_t>();
      dts[i] = t[1].cast<std::string>();
      shapes[i] = t[2].cast<std::vector<int64_t>>();
      a[i].dtype = dts[i].c_str();
      a[i].ndim = (int32_t)shapes[i].size();
      a[i].shape = shapes[i].data();
    }
    return iface_.launch_kernel(dev, (B200Stream)(uintptr_t)stream, op.c_str(), a.data(), num_inputs, (int32_t)args.size() - num_inputs) == 0;
  }

 private:
  struct Alloc { int device; size_t size; };
  void check(B200Status s, const char* what) const {
    if (s != 0) throw std::runtime_error("custom device '" + type_ + "': " + what + " failed with status " + std::to_string(s));
  }
  template <typename F> void need(F f, const char* what) const {
    if (!f) throw std::runtime_error("custom device '" + type_ + "': the plug-in does not implement " + what);
  }
  void bounds(uint64_t ptr, size_t n) {       // a copy must stay inside one live allocation
    std::lock_guard<std::mutex> g(mu_);
    auto it = live_.upper_bound((uintptr_t)ptr);
    if (it == live_.begin()) throw std::runtime_error("custom device: pointer is not inside a live allocation");
    --it;
    if ((uintptr_t)ptr + n > it->first + it->second.size) throw std::runtime_error("custom device: copy runs past the end of the allocation");
  }

  std::string path_, type_;
  void* handle_ = nullptr;
  B200DeviceInterface iface_;
  std::mutex mu_;
  std::map<uintptr_t, Alloc> live_;
  size_t allocated_ = 0, peak_ = 0;
};

void bind_custom_device(py::module_& m) {
  py::class_<CustomDevice, std::shared_ptr<CustomDevice>>(m, "CustomDevice")
      .def(py::init<const std::string&>())
      .def_property_readonly("device_type", &CustomDevice::type)
      .def_property_readonly("path", &CustomDevice::path)
      .def("device_count", &CustomDevice::device_count)
      .def("set_device", &CustomDevice::set_device)
      .def("malloc", &CustomDevice::malloc)
      .def("free", &CustomDevice::free)
      .def("memcpy_h2d", &CustomDevice::h2d)
      .def("memcpy_d2h", &CustomDevice::d2h)
      .def("memcpy_d2d", &CustomDevice::d2d)
      .def("memory_stats", &CustomDevice::memory_stats)
      .def("create_stream", &CustomDevice::create_stream)
      .def("destroy_stream", &CustomDevice::destroy_stream)
      .def("synchronize_stream", &CustomDevice::sync_stream)
      .def("create_event", &CustomDevice::create_event)
      .def("record_event", &CustomDevice::record_event)
      .def("synchronize_event", &CustomDevice::sync_event)
      .def("destroy_event", &CustomDevice::destroy_event)
      .def("synchronize", &CustomDevice::synchronize)
      .def("launch", &CustomDevice::launch);
}

}  // namespace runtime
}  // namespace b200
