// Native runtime pieces (symmetric peer heap, pinned staging / loader, graph executor). Bound into the _C module.
#pragma once
#include <torch/extension.h>

namespace b200 {
namespace runtime {
void bind(pybind11::module_& m);
void bind_symm(pybind11::module_& m);
void bind_loader(pybind11::module_& m);
void bind_graph(pybind11::module_& m);
void bind_tracer(pybind11::module_& m);
void bind_data_feed(pybind11::module_& m);
void bind_ir(pybind11::module_& m);
void bind_vision(pybind11::module_& m);
void bind_allocator(pybind11::module_& m);
void bind_custom_device(pybind11::module_& m);
}  // namespace runtime
}  // namespace b200
