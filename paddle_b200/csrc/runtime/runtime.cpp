#include "runtime.h"

namespace b200 {
namespace runtime {
void bind(pybind11::module_& m) {
  bind_symm(m);
  bind_loader(m);
  bind_graph(m);
  bind_tracer(m);
  bind_data_feed(m);
  bind_ir(m);
  bind_allocator(m);
  bind_custom_device(m);
  bind_vision(m);
}
}  // namespace runtime
}  // namespace b200
