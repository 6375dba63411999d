// Native program executor: dependency analysis + multi-stream scheduling + CUDA-graph instantiation.
// Parity (role): paddle/fluid/framework/new_executor (StandaloneExecutor / PirInterpreter: instruction list,
// dependency builder, stream analyzer inserting events between streams, GC).  Design: nodes are opaque callables
// (each launches kernels on the *current* stream); the executor topologically orders them, assigns each to its
// stream, inserts cudaEvent record/wait pairs on cross-stream edges only, and can capture one whole run into a
// CUDA graph that is replayed with a single launch.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <map>
#include <queue>
#include <stdexcept>
#include <vector>

#include "runtime.h"

namespace b200 {
namespace runtime {

class GraphExecutor {
 public:
  struct Node {
    pybind11::object fn;
    std::vector<int> deps;
    int stream = 0;
    int priority = 0;
    std::vector<int> wait_events;  // node ids whose completion event must be waited (cross-stream producers)
    bool record = false;           // some consumer on another stream needs our event
  };

  int add_node(pybind11::object fn, std::vector<int> deps, int stream, int priority) {
    Node n;
    n.fn = std::move(fn);
    n.deps = std::move(deps);
    n.stream = stream;
    n.priority = priority;
    for (int d : n.deps)
      if (d < 0 || d >= (int)nodes_.size()) throw std::runtime_error("GraphExecutor: dependency on unknown node");
    nodes_.push_back(std::move(n));
    finalized_ = false;
    return (int)nodes_.size() - 1;
  }

  // Kahn topological order, ties broken by (priority desc, insertion order); computes the minimal cross-stream event set.
  void finalize() {
    const int n = (int)nodes_.size();
    std::vector<int> indeg(n, 0);
    std::vector<std::vector<int>> out(n);
    for (int i = 0; i < n; ++i)
      for (int d : nodes_[i].deps) { out[d].push_back(i); ++indeg[i]; }
    auto cmp = [&](int a, int b) {
      if (nodes_[a].priority != nodes_[b].priority) return nodes_[a].priority < nodes_[b].priority;
      return a > b;
    };
    std::priority_queue<int, std::vector<int>, decltype(cmp)> ready(cmp);
    for (int i = 0; i < n; ++i) if (indeg[i] == 0) ready.push(i);
    order_.clear();
    while (!ready.empty()) {
      int u = ready.top(); ready.pop();
      order_.push_back(u);
      for (int v : out[u]) if (--indeg[v] == 0) ready.push(v);
    }
    if ((int)order_.size() != n) throw std::runtime_error("GraphExecutor: dependency cycle");
    // same-stream edges are ordered by the stream itself; for cross-stream edges keep only the latest producer per stream
    std::vector<int> pos(n);
    for (int i = 0; i < n; ++i) pos[order_[i]] = i;
    num_streams_ = 1;
    for (auto& nd : nodes_) { nd.wait_events.clear(); nd.record = false; num_streams_ = std::max(num_streams_, nd.stream + 1); }
    for (int i = 0; i < n; ++i) {
      std::map<int, int> latest;  // stream -> producer with max position
      for (int d : nodes_[i].deps) {
        if (nodes_[d].stream == nodes_[i].stream) continue;
        auto it = latest.find(nodes_[d].stream);
        if (it == latest.end() || pos[d] > pos[it->second]) latest[nodes_[d].stream] = d;
      }
      for (auto& kv : latest) { nodes_[i].wait_events.push_back(kv.second); nodes_[kv.second].record = true; }
    }
    finalized_ = true;
  }

  std::vector<int> order() { if (!finalized_) finalize(); return order_; }
  int num_cross_stream_edges() {
    if (!finalized_) finalize();
    int c = 0;
    for (auto& nd : nodes_) c += (int)nd.wait_events.size();
    return c;
  }

  // Eager run. On CPU-only builds every node simply runs in order.
  void run(int device) {
    if (!finalized_) finalize();
    const bool cuda = device >= 0 && at::cuda::is_available();
    if (!cuda) {
      for (int u : order_) nodes_[u].fn();
      return;
    }
    c10::cuda::CUDAGuard g(device);
    ensure_streams(device);
    auto origin = at::cuda::getCurrentCUDAStream(device);
    // side streams first wait for everything already queued on the caller's stream
    cudaEvent_t start = get_event(-1);
    cudaEventRecord(start, origin.stream());
    for (int s = 1; s < num_streams_; ++s) cudaStreamWaitEvent(streams_[s].stream(), start, 0);
    for (int u : order_) {
      Node& nd = nodes_[u];
      auto st = nd.stream == 0 ? origin : streams_[nd.stream];
      for (int p : nd.wait_events) cudaStreamWaitEvent(st.stream(), get_event(p), 0);
      {
        c10::cuda::CUDAStreamGuard sg(st);
        nd.fn();
      }
      if (nd.record) cudaEventRecord(get_event(u), st.stream());
    }
    // join side streams back into the caller's stream
    for (int s = 1; s < num_streams_; ++s) {
      cudaEvent_t e = get_event(-2 - s);
      cudaEventRecord(e, streams_[s].stream());
      cudaStreamWaitEvent(origin.stream(), e, 0);
    }
  }

  // Capture one run into a CUDA graph (all node callables must be capture-safe: no syncs, no allocations that escape).
  void capture(int device) {
    if (!at::cuda::is_available()) throw std::runtime_error("GraphExecutor.capture needs a GPU");
    c10::cuda::CUDAGuard g(device);
    auto cap = at::cuda::getStreamFromPool(false, device);
    cudaGraph_t graph = nullptr;
    {
      c10::cuda::CUDAStreamGuard sg(cap);
      if (cudaStreamBeginCapture(cap.stream(), cudaStreamCaptureModeThreadLocal) != cudaSuccess) throw std::runtime_error("BeginCapture failed");
      try { run(device); } catch (...) { cudaStreamEndCapture(cap.stream(), &graph); if (graph) cudaGraphDestroy(graph); throw; }
      if (cudaStreamEndCapture(cap.stream(), &graph) != cudaSuccess) throw std::runtime_error("EndCapture failed");
    }
    if (exec_) { cudaGraphExecDestroy(exec_); exec_ = nullptr; }
    if (cudaGraphInstantiate(&exec_, graph, 0) != cudaSuccess) { cudaGraphDestroy(graph); throw std::runtime_error("GraphInstantiate failed"); }
    size_t nn = 0;
    cudaGraphGetNodes(graph, nullptr, &nn);
    graph_nodes_ = (int)nn;
    cudaGraphDestroy(graph);
  }
  void replay(int device) {
    if (!exec_) throw std::runtime_error("GraphExecutor.replay before capture");
    c10::cuda::CUDAGuard g(device);
    if (cudaGraphLaunch(exec_, at::cuda::getCurrentCUDAStream(device).stream()) != cudaSuccess) throw std::runtime_error("GraphLaunch failed");
  }
  int graph_nodes() const { return graph_nodes_; }
  int num_nodes() const { return (int)nodes_.size(); }

  ~GraphExecutor() {
    if (exec_) cudaGraphExecDestroy(exec_);
    for (auto& kv : events_) cudaEventDestroy(kv.second);
  }

 private:
  void ensure_streams(int device) {
    while ((int)streams_.size() < num_streams_) streams_.push_back(at::cuda::getStreamFromPool(false, device));
  }
  cudaEvent_t get_event(int key) {
    auto it = events_.find(key);
    if (it != events_.end()) return it->second;
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    events_[key] = e;
    return e;
  }
  std::vector<Node> nodes_;
  std::vector<int> order_;
  bool finalized_ = false;
  int num_streams_ = 1;
  std::vector<c10::cuda::CUDAStream> streams_;
  std::map<int, cudaEvent_t> events_;
  cudaGraphExec_t exec_ = nullptr;
  int graph_nodes_ = 0;
};

void bind_graph(pybind11::module_& m) {
  pybind11::class_<GraphExecutor, std::shared_ptr<GraphExecutor>>(m, "GraphExecutor")
      .def(pybind11::init<>())
      .def("add_node", &GraphExecutor::add_node, pybind11::arg("fn"), pybind11::arg("deps") = std::vector<int>{},
           pybind11::arg("stream") = 0, pybind11::arg("priority") = 0)
      .def("finalize", &GraphExecutor::finalize)
      .def("order", &GraphExecutor::order)
      .def("num_cross_stream_edges", &GraphExecutor::num_cross_stream_edges)
      .def("run", &GraphExecutor::run, pybind11::arg("device") = -1)
      .def("capture", &GraphExecutor::capture)
      .def("replay", &GraphExecutor::replay)
      .def("graph_nodes", &GraphExecutor::graph_nodes)
      .def("num_nodes", &GraphExecutor::num_nodes);
}

}  // namespace runtime
}  // namespace b200
