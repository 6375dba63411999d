// A small SSA program IR with a pass manager: the native middle layer between a recorded static Program and its executor.
//
// Parity (role): paddle/pir (Program / Block / Operation / Value / Attribute, IrPrinter + parser, PassManager, pattern rewriter / DRR,
// DCE + CSE + constant folding + identity elimination transforms under paddle/fluid/pir/transforms, the inplace pass and the memory
// optimisation analysis of the new executor).  Design: values and operations live in flat tables owned by the Program (ids are stable,
// erasure is a tombstone, `compact()` renumbers); regions are nested Programs so control-flow ops carry their bodies; attributes are a
// closed variant (int / float / bool / string / int list / float list); every pass is a function Program& -> PassResult registered by
// name, so Python can compose pipelines and add declarative rewrite patterns ("DRR": a source DAG of op names + a replacement DAG)
// without recompiling.  Constant folding calls back into Python for the arithmetic (the kernels live there).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>

#include "runtime.h"

namespace py = pybind11;

namespace b200 {
namespace runtime {
namespace ir {

using Attr = std::variant<int64_t, double, bool, std::string, std::vector<int64_t>, std::vector<double>>;

struct Type {
  std::string dtype = "float32";
  std::vector<int64_t> shape;      // -1: dynamic
  bool operator==(const Type& o) const { return dtype == o.dtype && shape == o.shape; }
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) { if (d < 0) return -1; n *= d; }
    return n;
  }
};

static int64_t dtype_bytes(const std::string& d) {
  if (d == "float64" || d == "int64" || d == "complex64") return 8;
  if (d == "float32" || d == "int32") return 4;
  if (d == "float16" || d == "bfloat16" || d == "int16") return 2;
  if (d == "complex128") return 16;
  return 1;
}

struct Program;

struct Value {
  int id = -1;
  Type type;
  int def_op = -1;          // -1: block argument (input / parameter)
  std::string name;         // block arguments only
  std::string kind;         // "input" | "param" for block arguments
};

struct Op {
  int id = -1;
  std::string name;
  std::vector<int> operands;
  std::vector<int> results;
  std::map<std::string, Attr> attrs;
  std::vector<std::shared_ptr<Program>> regions;
  bool erased = false;
};

static const std::set<std::string>& impure_names() {
  static const std::set<std::string> s = {"dropout", "rand", "randn", "randint", "rand_like", "randn_like", "bernoulli", "multinomial", "normal", "uniform",
                                          "print", "assign", "set_value", "fill_", "zero_", "copy_", "all_reduce", "all_gather", "reduce_scatter", "send", "recv",
                                          "barrier", "py_func", "train_step", "backward", "while", "if", "feed", "fetch"};
  return s;
}

struct Program {
  std::vector<Value> values;
  std::vector<Op> ops;            // table indexed by op id
  std::vector<int> order;         // program order (op ids); rewrites insert here, ids stay stable
  int insert_at = -1;             // add_op position in `order` (-1: append)
  std::vector<int> outputs;       // values kept alive (fetch targets)
  std::map<int, int> replaced;    // rewrite log: value -> the value that took over its uses (front ends re-map fetch targets with it)

  int add_arg(const std::string& kind, const std::string& name, const Type& t) {
    Value v;
    v.id = (int)values.size(); v.type = t; v.name = name; v.kind = kind;
    values.push_back(v);
    return v.id;
  }
  std::vector<int> add_op(const std::string& name, const std::vector<int>& operands, const std::map<std::string, Attr>& attrs, const std::vector<Type>& result_types) {
    for (int o : operands)
      if (o < 0 || o >= (int)values.size()) throw std::runtime_error("ir: operand %" + std::to_string(o) + " of '" + name + "' is not defined");
    Op op;
    op.id = (int)ops.size(); op.name = name; op.operands = operands; op.attrs = attrs;
    for (const auto& t : result_types) {
      Value v;
      v.id = (int)values.size(); v.type = t; v.def_op = op.id;
      values.push_back(v);
      op.results.push_back(v.id);
    }
    ops.push_back(op);
    if (insert_at < 0) order.push_back(op.id);
    else order.insert(order.begin() + insert_at++, op.id);
    return ops.back().results;
  }
  bool is_pure(const Op& op) const {
    if (!op.regions.empty()) return false;
    const std::string& n = op.name;
    const std::string base = n.rfind("pd_op.", 0) == 0 ? n.substr(6) : n;
    if (!base.empty() && base.back() == '_') return false;          // in-place convention
    if (impure_names().count(base)) return false;
    auto it = op.attrs.find("side_effect");
    if (it != op.attrs.end() && std::holds_alternative<bool>(it->second) && std::get<bool>(it->second)) return false;
    return true;
  }
  std::vector<int> use_counts() const {
    std::vector<int> uses(values.size(), 0);
    for (const auto& op : ops)
      if (!op.erased)
        for (int o : op.operands) ++uses[o];
    for (int o : outputs) ++uses[o];
    return uses;
  }
  void replace_all_uses(int from, int to) {
    replaced[from] = to;
    for (auto& op : ops)
      if (!op.erased)
        for (auto& o : op.operands)
          if (o == from) o = to;
    for (auto& o : outputs)
      if (o == from) o = to;
  }
  int live_ops() const {
    int n = 0;
    for (const auto& op : ops) n += op.erased ? 0 : 1;
    return n;
  }
  // SSA dominance (definitions precede uses in program order), result / operand bookkeeping
  void verify() const {
    std::vector<char> defined(values.size(), 0);
    for (const auto& v : values)
      if (v.def_op < 0) defined[v.id] = 1;
    for (int oid : order) {
      const Op& op = ops[oid];
      if (op.erased) continue;
      for (int o : op.operands) {
        if (o < 0 || o >= (int)values.size()) throw std::runtime_error("ir verify: '" + op.name + "' uses an unknown value");
        if (!defined[o]) throw std::runtime_error("ir verify: '" + op.name + "' uses %" + std::to_string(o) + " before its definition");
      }
      for (int r : op.results) {
        if (values[r].def_op != op.id) throw std::runtime_error("ir verify: result %" + std::to_string(r) + " does not point back to '" + op.name + "'");
        defined[r] = 1;
      }
      for (const auto& reg : op.regions) reg->verify();
    }
    for (int o : outputs)
      if (o < 0 || o >= (int)values.size() || !defined[o]) throw std::runtime_error("ir verify: program output %" + std::to_string(o) + " is not defined");
  }
};

// Text is assembled with this appender instead of Out: inserting a number into an iostream looks up the num_put facet of the
// global locale, which segfaults in this process (the extension is compiled against another libstdc++ than the one loaded first).
struct Out {
  std::string s;
  Out& operator<<(const std::string& v) { s += v; return *this; }
  Out& operator<<(const char* v) { s += v; return *this; }
  Out& operator<<(char v) { s += v; return *this; }
  Out& operator<<(int v) { s += std::to_string(v); return *this; }
  Out& operator<<(int64_t v) { s += std::to_string(v); return *this; }
  Out& operator<<(size_t v) { s += std::to_string(v); return *this; }
  Out& operator<<(double v) { char buf[40]; snprintf(buf, sizeof(buf), "%.17g", v); s += buf; return *this; }
  const std::string& str() const { return s; }
};

// ------------------------------------------------------------------------------------------------ printer / parser
static std::string type_str(const Type& t) {
  Out os;
  os << "tensor<";
  for (auto d : t.shape) os << (d < 0 ? std::string("?") : std::to_string(d)) << "x";
  os << t.dtype << ">";
  return os.str();
}
static std::string quote(const std::string& s) {
  std::string o = "\"";
  for (char c : s) { if (c == '"' || c == '\\') o += '\\'; o += c; }
  return o + "\"";
}
static std::string attr_str(const Attr& a) {
  Out os;
  if (auto p = std::get_if<int64_t>(&a)) os << *p;
  else if (auto p = std::get_if<double>(&a)) { os << *p; if (os.str().find_first_of(".enai") == std::string::npos) os << ".0"; }
  else if (auto p = std::get_if<bool>(&a)) os << (*p ? "true" : "false");
  else if (auto p = std::get_if<std::string>(&a)) os << quote(*p);
  else if (auto p = std::get_if<std::vector<int64_t>>(&a)) { os << "["; for (size_t i = 0; i < p->size(); ++i) os << (i ? ", " : "") << (*p)[i]; os << "]"; }
  else if (auto p = std::get_if<std::vector<double>>(&a)) { os << "[f "; for (size_t i = 0; i < p->size(); ++i) os << (i ? ", " : "") << (*p)[i]; os << "]"; }
  return os.str();
}
static void print_program(const Program& p, Out& os, int indent) {
  const std::string pad(indent, ' ');
  os << pad << "program {\n";
  for (const auto& v : p.values)
    if (v.def_op < 0) os << pad << "  %" << v.id << " = " << v.kind << " " << quote(v.name) << " : " << type_str(v.type) << "\n";
  for (int oid : p.order) {
    const Op& op = p.ops[oid];
    if (op.erased) continue;
    os << pad << "  ";
    for (size_t i = 0; i < op.results.size(); ++i) os << (i ? ", " : "") << "%" << op.results[i];
    if (!op.results.empty()) os << " = ";
    os << op.name << "(";
    for (size_t i = 0; i < op.operands.size(); ++i) os << (i ? ", " : "") << "%" << op.operands[i];
    os << ")";
    if (!op.attrs.empty()) {
      os << " {";
      bool first = true;
      for (const auto& kv : op.attrs) { os << (first ? "" : ", ") << kv.first << " = " << attr_str(kv.second); first = false; }
      os << "}";
    }
    if (!op.results.empty()) {
      os << " : ";
      for (size_t i = 0; i < op.results.size(); ++i) os << (i ? ", " : "") << type_str(p.values[op.results[i]].type);
    }
    os << "\n";
    for (const auto& r : op.regions) print_program(*r, os, indent + 4);
  }
  os << pad << "  return";
  for (size_t i = 0; i < p.outputs.size(); ++i) os << (i ? "," : "") << " %" << p.outputs[i];
  os << "\n" << pad << "}\n";
}

struct Parser {
  const std::string& s;
  size_t i = 0;
  explicit Parser(const std::string& src) : s(src) {}
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i; }
  bool eat(const std::string& t) { ws(); if (s.compare(i, t.size(), t) == 0) { i += t.size(); return true; } return false; }
  void expect(const std::string& t) { if (!eat(t)) fail("expected '" + t + "'"); }
  [[noreturn]] void fail(const std::string& m) { throw std::runtime_error("ir parse error at offset " + std::to_string(i) + ": " + m); }
  std::string ident() {
    ws();
    size_t b = i;
    while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '.')) ++i;
    if (b == i) fail("identifier expected");
    return s.substr(b, i - b);
  }
  int value_ref() { expect("%"); size_t b = i; while (i < s.size() && isdigit((unsigned char)s[i])) ++i; if (b == i) fail("value id expected"); return std::stoi(s.substr(b, i - b)); }
  std::string str() {
    expect("\"");
    std::string o;
    while (i < s.size() && s[i] != '"') { if (s[i] == '\\') ++i; o += s[i++]; }
    expect("\"");
    return o;
  }
  Type type() {
    expect("tensor<");
    Type t;
    while (true) {
      ws();
      size_t b = i;
      if (s[i] == '?') { ++i; if (i < s.size() && s[i] == 'x') { ++i; t.shape.push_back(-1); continue; } fail("bad dynamic dim"); }
      while (i < s.size() && isdigit((unsigned char)s[i])) ++i;
      if (i > b && i < s.size() && s[i] == 'x') { t.shape.push_back(std::stoll(s.substr(b, i - b))); ++i; continue; }
      i = b;
      break;
    }
    t.dtype = ident();
    expect(">");
    return t;
  }
  Attr attr() {
    ws();
    if (s[i] == '"') return str();
    if (eat("true")) return true;
    if (eat("false")) return false;
    if (s[i] == '[') {
      ++i;
      const bool fl = eat("f ");
      std::vector<int64_t> vi;
      std::vector<double> vf;
      ws();
      while (s[i] != ']') {
        size_t b = i;
        while (i < s.size() && s[i] != ',' && s[i] != ']') ++i;
        const std::string tok = s.substr(b, i - b);
        if (fl) vf.push_back(std::stod(tok)); else vi.push_back(std::stoll(tok));
        if (s[i] == ',') ++i;
        ws();
      }
      ++i;
      if (fl) return vf;
      return vi;
    }
    size_t b = i;
    while (i < s.size() && (isdigit((unsigned char)s[i]) || strchr("+-.einfa", s[i]))) ++i;
    const std::string tok = s.substr(b, i - b);
    if (tok.empty()) fail("attribute value expected");
    if (tok.find_first_of(".enai") != std::string::npos) return std::stod(tok);
    return (int64_t)std::stoll(tok);
  }
  std::shared_ptr<Program> program() {
    auto p = std::make_shared<Program>();
    std::map<int, int> remap;     // printed id -> new id
    auto use = [&](int printed) {
      auto it = remap.find(printed);
      if (it == remap.end()) fail("%" + std::to_string(printed) + " is used before its definition");
      return it->second;
    };
    expect("program");
    expect("{");
    while (true) {
      ws();
      if (eat("return")) {
        ws();
        while (i < s.size() && s[i] == '%') { p->outputs.push_back(use(value_ref())); eat(","); ws(); }
        expect("}");
        return p;
      }
      std::vector<int> res;
      size_t save = i;
      if (s[i] == '%') {
        while (true) { res.push_back(value_ref()); if (!eat(",")) break; }
        expect("=");
      } else {
        i = save;
      }
      const std::string name = ident();
      if ((name == "input" || name == "param") && res.size() == 1) {
        const std::string n = str();
        expect(":");
        remap[res[0]] = p->add_arg(name, n, type());
        continue;
      }
      expect("(");
      std::vector<int> operands;
      ws();
      while (s[i] == '%') { operands.push_back(use(value_ref())); eat(","); ws(); }
      expect(")");
      std::map<std::string, Attr> attrs;
      if (eat("{")) {
        ws();
        while (s[i] != '}') { const std::string k = ident(); expect("="); attrs[k] = attr(); eat(","); ws(); }
        ++i;
      }
      std::vector<Type> types;
      if (!res.empty()) { expect(":"); while (true) { types.push_back(type()); if (!eat(",")) break; } }
      if (types.size() != res.size()) fail("result / type count mismatch for '" + name + "'");
      auto ids = p->add_op(name, operands, attrs, types);
      for (size_t k = 0; k < res.size(); ++k) remap[res[k]] = ids[k];
      ws();
      while (s.compare(i, 7, "program") == 0) { p->ops.back().regions.push_back(program()); ws(); }
    }
  }
};

// ------------------------------------------------------------------------------------------------ passes
struct PassResult {
  std::string name;
  int ops_before = 0, ops_after = 0, changed = 0;
};

static int pass_dce(Program& p) {
  int removed = 0;
  bool again = true;
  while (again) {
    again = false;
    auto uses = p.use_counts();
    for (auto it = p.order.rbegin(); it != p.order.rend(); ++it) {
      Op& op = p.ops[*it];
      if (op.erased || !p.is_pure(op)) continue;
      bool live = false;
      for (int r : op.results) live |= uses[r] > 0;
      if (!live) {
        op.erased = true;
        for (int o : op.operands) --uses[o];
        ++removed;
        again = true;
      }
    }
  }
  for (auto& op : p.ops)
    if (!op.erased)
      for (auto& r : op.regions) removed += pass_dce(*r);
  return removed;
}

static std::string op_key(const Op& op) {
  Out os;
  os << op.name << "(";
  for (int o : op.operands) os << o << ",";
  os << "){";
  for (const auto& kv : op.attrs) os << kv.first << "=" << attr_str(kv.second) << ";";
  os << "}" << op.results.size();
  return os.str();
}

static int pass_cse(Program& p) {
  int merged = 0;
  std::unordered_map<std::string, int> seen;      // key -> op id
  for (int oid : p.order) {
    Op& op = p.ops[oid];
    if (op.erased || !p.is_pure(op) || op.results.empty()) continue;
    const std::string key = op_key(op);
    auto it = seen.find(key);
    if (it == seen.end()) { seen.emplace(key, op.id); continue; }
    const Op& first = p.ops[it->second];
    bool same_types = first.results.size() == op.results.size();
    for (size_t k = 0; same_types && k < op.results.size(); ++k) same_types = p.values[first.results[k]].type == p.values[op.results[k]].type;
    if (!same_types) continue;
    for (size_t k = 0; k < op.results.size(); ++k) p.replace_all_uses(op.results[k], first.results[k]);
    op.erased = true;
    ++merged;
  }
  return merged;
}

static const Attr* find_attr(const Op& op, const std::string& k) {
  auto it = op.attrs.find(k);
  return it == op.attrs.end() ? nullptr : &it->second;
}
static std::string base_name(const std::string& n) { return n.rfind("pd_op.", 0) == 0 ? n.substr(6) : n; }

// reshape / cast / transpose / scale that do nothing, transpose o transpose = identity, reshape o reshape = the outer reshape
static int pass_identity_elim(Program& p) {
  int changed = 0;
  for (size_t oi = 0; oi < p.order.size(); ++oi) {
    Op& op = p.ops[p.order[oi]];
    if (op.erased || op.operands.empty() || op.results.size() != 1) continue;
    const std::string n = base_name(op.name);
    const Type& in = p.values[op.operands[0]].type;
    const Type& out = p.values[op.results[0]].type;
    bool identity = false;
    if ((n == "reshape" || n == "view" || n == "flatten" || n == "squeeze" || n == "unsqueeze" || n == "expand" || n == "contiguous" || n == "clone_view") && op.operands.size() == 1)
      identity = in == out && in.numel() >= 0;
    else if ((n == "cast" || n == "to" || n == "astype") && op.operands.size() == 1)
      identity = in == out;
    else if ((n == "transpose" || n == "permute") && op.operands.size() == 1) {
      if (auto perm = find_attr(op, "perm"))
        if (auto v = std::get_if<std::vector<int64_t>>(perm)) {
          identity = true;
          for (size_t i = 0; i < v->size(); ++i) identity &= (*v)[i] == (int64_t)i;
        }
    } else if (n == "scale" && op.operands.size() == 1) {
      const Attr* sc = find_attr(op, "scale");
      const Attr* bi = find_attr(op, "bias");
      identity = sc && std::holds_alternative<double>(*sc) && std::get<double>(*sc) == 1.0 && (!bi || (std::holds_alternative<double>(*bi) && std::get<double>(*bi) == 0.0));
    }
    else if ((n == "dropout" || n == "dropout2d" || n == "dropout3d" || n == "alpha_dropout" || n == "feature_alpha_dropout") && op.operands.size() == 1 && in == out) {
      // inference clone / eval mode: dropout(training=False) in the default upscale_in_train mode, or p = 0, passes its input through
      auto is_false = [&](const char* k) { const Attr* a = find_attr(op, k); return a && std::holds_alternative<bool>(*a) && !std::get<bool>(*a); };
      auto is_zero = [&](const char* k) { const Attr* a = find_attr(op, k); return a && ((std::holds_alternative<double>(*a) && std::get<double>(*a) == 0.0) || (std::holds_alternative<int64_t>(*a) && std::get<int64_t>(*a) == 0)); };
      auto downscale = [&](const char* k) { const Attr* a = find_attr(op, k); return a && std::holds_alternative<std::string>(*a) && std::get<std::string>(*a) == "downscale_in_infer"; };
      const bool eval = is_false("ktraining") || is_false("a3");
      identity = (is_zero("a1") || is_zero("kp")) || (eval && !downscale("a4") && !downscale("kmode"));
    }
    if (identity) {
      p.replace_all_uses(op.results[0], op.operands[0]);
      op.erased = true;
      ++changed;
      continue;
    }
    // producer-consumer folds
    const int src = op.operands[0];
    const int def = p.values[src].def_op;
    if (def < 0 || p.ops[def].erased || p.ops[def].operands.size() != 1) continue;
    Op& prod = p.ops[def];
    const std::string pn = base_name(prod.name);
    if ((n == "transpose" || n == "permute") && (pn == "transpose" || pn == "permute")) {
      const Attr* a = find_attr(op, "perm");
      const Attr* b = find_attr(prod, "perm");
      if (a && b && std::holds_alternative<std::vector<int64_t>>(*a) && std::holds_alternative<std::vector<int64_t>>(*b)) {
        const auto& outer = std::get<std::vector<int64_t>>(*a);
        const auto& inner = std::get<std::vector<int64_t>>(*b);
        if (outer.size() == inner.size()) {
          std::vector<int64_t> comp(outer.size());
          bool id = true;
          for (size_t i = 0; i < outer.size(); ++i) { comp[i] = inner[outer[i]]; id &= comp[i] == (int64_t)i; }
          if (id) {
            p.replace_all_uses(op.results[0], prod.operands[0]);
            op.erased = true;
          } else {
            op.operands[0] = prod.operands[0];
            op.attrs["perm"] = comp;
          }
          ++changed;
        }
      }
    } else if ((n == "reshape" || n == "view") && (pn == "reshape" || pn == "view") && op.operands.size() == 1) {
      op.operands[0] = prod.operands[0];      // the inner reshape may become dead: DCE removes it
      ++changed;
    }
  }
  return changed;
}

// ---- declarative rewrite patterns ("DRR").  Source: a DAG of ops ending in the anchor; inner results must have no other use.
struct PatOp {
  std::string name;
  std::vector<std::string> ins;
  std::vector<std::string> outs;
  std::map<std::string, Attr> attrs;      // source: constraints; result: attributes to set ("$sym.attr" strings copy from a matched op)
};
struct Pattern {
  std::string name;
  std::vector<PatOp> source;              // topological, last = anchor
  std::vector<PatOp> result;
};

static bool match_op(const Program& p, const std::vector<int>& uses, const Pattern& pat, int pi, int op_id, std::map<std::string, int>& sym, std::vector<int>& matched,
                     const std::set<std::string>& inner_syms) {
  const PatOp& po = pat.source[pi];
  const Op& op = p.ops[op_id];
  if (op.erased || base_name(op.name) != po.name || op.operands.size() != po.ins.size() || op.results.size() != po.outs.size() || !op.regions.empty()) return false;
  for (const auto& kv : po.attrs) {
    const Attr* a = find_attr(op, kv.first);
    if (!a || !(*a == kv.second)) return false;
  }
  for (size_t k = 0; k < po.outs.size(); ++k) {
    auto it = sym.find(po.outs[k]);
    if (it != sym.end() && it->second != op.results[k]) return false;
    sym[po.outs[k]] = op.results[k];
  }
  matched[pi] = op_id;
  for (size_t k = 0; k < po.ins.size(); ++k) {
    const std::string& s = po.ins[k];
    const int v = op.operands[k];
    if (inner_syms.count(s)) {
      // produced by an earlier pattern op: find it, require single use, recurse
      int producer = -1;
      for (int q = 0; q < pi; ++q)
        for (const auto& o : pat.source[q].outs)
          if (o == s) producer = q;
      const int def = p.values[v].def_op;
      if (def < 0 || uses[v] != 1) return false;
      if (matched[producer] >= 0) { if (matched[producer] != def) return false; continue; }
      if (!match_op(p, uses, pat, producer, def, sym, matched, inner_syms)) return false;
    } else {
      auto it = sym.find(s);
      if (it != sym.end() && it->second != v) return false;
      sym[s] = v;
    }
  }
  return true;
}

using InferFn = std::function<std::vector<Type>(const std::string&, const std::vector<Type>&)>;

static int apply_pattern(Program& p, const Pattern& pat, const InferFn& infer) {
  int applied = 0;
  std::set<std::string> inner;
  for (size_t q = 0; q + 1 < pat.source.size(); ++q)
    for (const auto& o : pat.source[q].outs) inner.insert(o);
  for (size_t pos = 0; pos < p.order.size(); ++pos) {
    const int anchor = p.order[pos];
    if (p.ops[anchor].erased) continue;
    auto uses = p.use_counts();
    std::map<std::string, int> sym;
    std::vector<int> matched(pat.source.size(), -1);
    if (!match_op(p, uses, pat, (int)pat.source.size() - 1, anchor, sym, matched, inner)) continue;
    bool complete = true;
    for (int m : matched) complete &= m >= 0;
    if (!complete) continue;
    // every external input must be defined before the anchor (it is: each feeds a matched op that precedes or is the anchor)
    const std::vector<int> anchor_results = p.ops[anchor].results;
    std::map<std::string, int> rsym = sym;
    const size_t table_before = p.ops.size(), values_before = p.values.size();
    p.insert_at = (int)pos;               // replacement ops take the anchor's place in program order
    bool ok = true;
    for (size_t ri = 0; ri < pat.result.size() && ok; ++ri) {
      const PatOp& ro = pat.result[ri];
      std::vector<int> operands;
      std::vector<Type> in_types;
      for (const auto& s : ro.ins) {
        auto it = rsym.find(s);
        if (it == rsym.end()) { ok = false; break; }
        operands.push_back(it->second);
        in_types.push_back(p.values[it->second].type);
      }
      if (!ok) break;
      std::map<std::string, Attr> attrs;
      for (const auto& kv : ro.attrs) {
        const std::string* sv = std::get_if<std::string>(&kv.second);
        if (sv && sv->size() > 1 && (*sv)[0] == '$' && sv->find('.') != std::string::npos) {      // "$sym.attr": copy from the op that produced sym
          const size_t dot = sv->find('.');
          const std::string out_sym = sv->substr(1, dot - 1), key = sv->substr(dot + 1);
          for (size_t q = 0; q < pat.source.size(); ++q)
            for (const auto& o : pat.source[q].outs)
              if (o == out_sym)
                if (const Attr* a = find_attr(p.ops[matched[q]], key)) attrs[kv.first] = *a;
        } else {
          attrs[kv.first] = kv.second;
        }
      }
      std::vector<Type> out_types;
      if (ri + 1 == pat.result.size()) {
        for (int r : anchor_results) out_types.push_back(p.values[r].type);
      } else {
        out_types = infer ? infer(ro.name, in_types) : std::vector<Type>();
        if (out_types.size() != ro.outs.size()) { ok = false; break; }
      }
      auto ids = p.add_op(ro.name, operands, attrs, out_types);
      for (size_t k = 0; k < ro.outs.size() && k < ids.size(); ++k) rsym[ro.outs[k]] = ids[k];
    }
    const int inserted = p.insert_at - (int)pos;
    p.insert_at = -1;
    if (!ok || p.ops.size() == table_before) {      // roll back a half-built replacement
      p.order.erase(p.order.begin() + pos, p.order.begin() + pos + inserted);
      p.ops.resize(table_before);
      p.values.resize(values_before);
      continue;
    }
    const Op& last_new = p.ops.back();
    for (size_t k = 0; k < anchor_results.size() && k < last_new.results.size(); ++k) p.replace_all_uses(anchor_results[k], last_new.results[k]);
    for (int m : matched) p.ops[m].erased = true;
    pos += inserted;                      // continue after the (now erased) anchor
    ++applied;
  }
  return applied;
}

// ---- analyses ------------------------------------------------------------------------------------------------------------------
// Liveness-based buffer plan for intermediates with static shapes: first-fit over [def, last use] intervals in program order.
struct MemoryPlan {
  std::map<int, int64_t> offset;     // value id -> byte offset in the workspace
  int64_t peak = 0, naive = 0;
};
static MemoryPlan plan_memory(const Program& p, int64_t align) {
  MemoryPlan plan;
  std::map<int, int> first, last;
  std::vector<int> live_order;
  for (size_t pos = 0; pos < p.order.size(); ++pos) {
    const Op& op = p.ops[p.order[pos]];
    if (op.erased) continue;
    for (int r : op.results) { first[r] = (int)pos; last[r] = (int)pos; live_order.push_back(r); }
    for (int o : op.operands) if (first.count(o)) last[o] = (int)pos;
  }
  for (int o : p.outputs) if (first.count(o)) last[o] = (int)p.order.size();
  struct Blk { int64_t off, size; int end; };
  std::vector<Blk> active;
  for (int v : live_order) {
    const Type& t = p.values[v].type;
    const int64_t n = t.numel();
    if (n < 0) continue;               // dynamic: not planned
    int64_t size = (n * dtype_bytes(t.dtype) + align - 1) / align * align;
    if (size == 0) size = align;
    plan.naive += size;
    active.erase(std::remove_if(active.begin(), active.end(), [&](const Blk& b) { return b.end < first[v]; }), active.end());
    std::sort(active.begin(), active.end(), [](const Blk& a, const Blk& b) { return a.off < b.off; });
    int64_t off = 0;
    for (const auto& b : active) {
      if (off + size <= b.off) break;
      off = std::max(off, b.off + b.size);
    }
    active.push_back({off, size, last[v]});
    plan.offset[v] = off;
    plan.peak = std::max(plan.peak, off + size);
  }
  return plan;
}

// elementwise ops whose first operand dies here and has the result's type may write in place
static int pass_inplace(Program& p) {
  static const std::set<std::string> ew = {"add", "sub", "subtract", "mul", "multiply", "div", "divide", "relu", "gelu", "silu", "sigmoid", "tanh", "exp", "scale", "neg",
                                           "abs", "sqrt", "rsqrt", "square", "clip", "clamp", "leaky_relu", "dropout", "cast", "where", "maximum", "minimum", "pow", "swiglu"};
  int marked = 0;
  std::map<int, int> last_use;
  for (size_t pos = 0; pos < p.order.size(); ++pos) {
    const Op& op = p.ops[p.order[pos]];
    if (op.erased) continue;
    for (int o : op.operands) last_use[o] = (int)pos;
  }
  for (int o : p.outputs) last_use[o] = (int)p.order.size();
  for (size_t pos = 0; pos < p.order.size(); ++pos) {
    Op& op = p.ops[p.order[pos]];
    if (op.erased || op.results.size() != 1 || op.operands.empty() || !ew.count(base_name(op.name))) continue;
    const int src = op.operands[0];
    if (p.values[src].def_op < 0) continue;                          // never overwrite inputs / parameters
    if (last_use[src] != (int)pos) continue;
    if (std::count(op.operands.begin(), op.operands.end(), src) != 1) continue;
    if (!(p.values[src].type == p.values[op.results[0]].type)) continue;
    op.attrs["inplace"] = true;
    ++marked;
  }
  return marked;
}

static void compact(Program& p) {
  // drop tombstones and renumber ops (values keep their ids: uses stay valid)
  std::vector<Op> ops;
  std::vector<int> order;
  std::map<int, int> remap;
  for (int oid : p.order) {
    Op& op = p.ops[oid];
    if (op.erased) continue;
    remap[oid] = (int)ops.size();
    Op c = op;
    c.id = (int)ops.size();
    ops.push_back(std::move(c));
    order.push_back(ops.back().id);
  }
  for (auto& v : p.values)
    if (v.def_op >= 0) { auto it = remap.find(v.def_op); v.def_op = it == remap.end() ? -2 : it->second; }   // -2: value of an erased op
  p.ops = std::move(ops);
  p.order = std::move(order);
}

// ------------------------------------------------------------------------------------------------ python surface
static Attr attr_from_py(const py::handle& h) {
  if (py::isinstance<py::bool_>(h)) return h.cast<bool>();
  if (py::isinstance<py::int_>(h)) return h.cast<int64_t>();
  if (py::isinstance<py::float_>(h)) return h.cast<double>();
  if (py::isinstance<py::str>(h)) return h.cast<std::string>();
  if (py::isinstance<py::list>(h) || py::isinstance<py::tuple>(h)) {
    bool all_int = true;
    for (auto e : h) all_int &= py::isinstance<py::int_>(e) && !py::isinstance<py::bool_>(e);
    if (all_int) { std::vector<int64_t> v; for (auto e : h) v.push_back(e.cast<int64_t>()); return v; }
    std::vector<double> v;
    for (auto e : h) v.push_back(e.cast<double>());
    return v;
  }
  throw std::runtime_error("ir: attribute values are int / float / bool / str / list of numbers");
}
static py::object attr_to_py(const Attr& a) {
  return std::visit([](auto&& v) -> py::object { return py::cast(v); }, a);
}
static std::map<std::string, Attr> attrs_from_py(const py::dict& d) {
  std::map<std::string, Attr> m;
  for (auto kv : d) m[kv.first.cast<std::string>()] = attr_from_py(kv.second);
  return m;
}
static Type type_from_py(const py::handle& h) {      // (dtype, shape)
  auto t = h.cast<py::tuple>();
  Type ty;
  ty.dtype = t[0].cast<std::string>();
  for (auto d : t[1]) ty.shape.push_back(d.cast<int64_t>());
  return ty;
}
static py::tuple type_to_py(const Type& t) { return py::make_tuple(t.dtype, py::cast(t.shape)); }

static PatOp patop_from_py(const py::handle& h) {    // (name, [ins], [outs], {attrs})
  auto t = h.cast<py::tuple>();
  PatOp po;
  po.name = t[0].cast<std::string>();
  po.ins = t[1].cast<std::vector<std::string>>();
  po.outs = t[2].cast<std::vector<std::string>>();
  if (t.size() > 3) po.attrs = attrs_from_py(t[3].cast<py::dict>());
  return po;
}

class PassManager {
 public:
  explicit PassManager(std::vector<std::string> passes) : passes_(std::move(passes)) {}
  void add_pass(const std::string& n) { passes_.push_back(n); }
  void add_pattern(const std::string& name, const py::list& source, const py::list& result) {
    Pattern p;
    p.name = name;
    for (auto h : source) p.source.push_back(patop_from_py(h));
    for (auto h : result) p.result.push_back(patop_from_py(h));
    if (p.source.empty() || p.result.empty()) throw std::runtime_error("ir: a rewrite pattern needs a source and a result");
    patterns_[name] = p;
  }
  void set_folder(py::object f) { folder_ = std::move(f); }
  void set_type_infer(py::object f) { infer_ = std::move(f); }
  void enable_ir_printing(bool on) { print_ = on; }

  py::list run(Program& prog) {
    py::list report;
    for (const auto& name : passes_) {
      PassResult r;
      r.name = name;
      r.ops_before = prog.live_ops();
      if (name == "dce") r.changed = pass_dce(prog);
      else if (name == "cse") r.changed = pass_cse(prog);
      else if (name == "identity_elim") r.changed = pass_identity_elim(prog);
      else if (name == "inplace") r.changed = pass_inplace(prog);
      else if (name == "constant_fold") r.changed = fold(prog);
      else if (name == "compact") { compact(prog); r.changed = 0; }
      else if (patterns_.count(name)) {
        InferFn infer;
        if (!infer_.is_none())
          infer = [this](const std::string& op, const std::vector<Type>& ins) {
            py::list l;
            for (const auto& t : ins) l.append(type_to_py(t));
            std::vector<Type> out;
            py::object res = infer_(op, l);
            if (res.is_none()) return out;
            for (auto h : res) out.push_back(type_from_py(h));
            return out;
          };
        r.changed = apply_pattern(prog, patterns_[name], infer);
      } else {
        throw std::runtime_error("ir: unknown pass '" + name + "'");
      }
      prog.verify();
      r.ops_after = prog.live_ops();
      py::dict d;
      d["pass"] = r.name; d["ops_before"] = r.ops_before; d["ops_after"] = r.ops_after; d["changed"] = r.changed;
      if (print_) { Out os; print_program(prog, os, 0); d["ir"] = os.str(); }
      report.append(d);
    }
    return report;
  }

 private:
  // ops whose operands are all `constant` ops: ask Python for the value (const ids index a Python-side table), replace by a constant
  int fold(Program& p) {
    if (folder_.is_none()) return 0;
    int folded = 0;
    for (size_t pos = 0; pos < p.order.size(); ++pos) {
      Op& op = p.ops[p.order[pos]];
      if (op.erased || !p.is_pure(op) || op.results.size() != 1 || op.operands.empty() || base_name(op.name) == "constant") continue;
      py::list ids;
      bool all_const = true;
      for (int o : op.operands) {
        const int def = p.values[o].def_op;
        if (def < 0 || p.ops[def].erased || base_name(p.ops[def].name) != "constant") { all_const = false; break; }
        const Attr* a = find_attr(p.ops[def], "const_id");
        if (!a || !std::holds_alternative<int64_t>(*a)) { all_const = false; break; }
        ids.append(std::get<int64_t>(*a));
      }
      if (!all_const) continue;
      py::dict attrs;
      for (const auto& kv : op.attrs) attrs[py::str(kv.first)] = attr_to_py(kv.second);
      py::object res = folder_(base_name(op.name), ids, attrs);
      if (res.is_none()) continue;
      const Type ty = p.values[op.results[0]].type;
      p.insert_at = (int)pos;
      auto ids_new = p.add_op("pd_op.constant", {}, {{"const_id", (int64_t)res.cast<int64_t>()}}, {ty});
      p.insert_at = -1;
      Op& old = p.ops[p.order[pos + 1]];       // `op` may dangle after the table grew: re-fetch
      p.replace_all_uses(old.results[0], ids_new[0]);
      old.erased = true;
      ++pos;
      ++folded;
    }
    return folded;
  }

  std::vector<std::string> passes_;
  std::map<std::string, Pattern> patterns_;
  py::object folder_ = py::none(), infer_ = py::none();
  bool print_ = false;
};

}  // namespace ir

void bind_ir(py::module_& m) {
  using namespace ir;
  py::class_<Program, std::shared_ptr<Program>>(m, "IrProgram")
      .def(py::init<>())
      .def("add_input", [](Program& p, const std::string& name, const std::string& dtype, const std::vector<int64_t>& shape) { return p.add_arg("input", name, Type{dtype, shape}); })
      .def("add_param", [](Program& p, const std::string& name, const std::string& dtype, const std::vector<int64_t>& shape) { return p.add_arg("param", name, Type{dtype, shape}); })
      .def("add_op", [](Program& p, const std::string& name, const std::vector<int>& operands, const py::dict& attrs, const py::list& result_types) {
        std::vector<Type> tys;
        for (auto h : result_types) tys.push_back(type_from_py(h));
        return p.add_op(name, operands, attrs_from_py(attrs), tys);
      }, py::arg("name"), py::arg("operands"), py::arg("attrs") = py::dict(), py::arg("result_types") = py::list())
      .def("add_region", [](Program& p, int op_id, std::shared_ptr<Program> body) {
        if (op_id < 0 || op_id >= (int)p.ops.size()) throw std::runtime_error("ir: unknown op");
        p.ops[op_id].regions.push_back(std::move(body));
      })
      .def("set_outputs", [](Program& p, const std::vector<int>& outs) { p.outputs = outs; })
      .def("outputs", [](const Program& p) { return p.outputs; })
      .def("num_ops", &Program::live_ops)
      .def("verify", &Program::verify)
      .def("value_type", [](const Program& p, int v) { return type_to_py(p.values.at(v).type); })
      .def("value_info", [](const Program& p, int v) {
        const Value& val = p.values.at(v);
        py::dict d;
        d["id"] = val.id; d["type"] = type_to_py(val.type); d["def_op"] = val.def_op; d["name"] = val.name; d["kind"] = val.kind;
        return d;
      })
      .def("args", [](const Program& p) {
        py::list l;
        for (const auto& v : p.values)
          if (v.def_op == -1) l.append(py::make_tuple(v.id, v.kind, v.name, type_to_py(v.type)));
        return l;
      })
      .def("ops", [](const Program& p) {
        py::list l;
        for (int oid : p.order) {
          const Op& op = p.ops[oid];
          if (op.erased) continue;
          py::dict d;
          py::dict attrs;
          for (const auto& kv : op.attrs) attrs[py::str(kv.first)] = attr_to_py(kv.second);
          d["id"] = op.id; d["name"] = op.name; d["operands"] = op.operands; d["results"] = op.results; d["attrs"] = attrs; d["num_regions"] = (int)op.regions.size();
          l.append(d);
        }
        return l;
      })
      .def("region", [](const Program& p, int op_id, int k) { return p.ops.at(op_id).regions.at(k); })
      .def("use_counts", &Program::use_counts)
      .def("replace_all_uses", &Program::replace_all_uses)
      .def("replacements", [](const Program& p) { return p.replaced; })
      .def("erase_op", [](Program& p, int op_id) { p.ops.at(op_id).erased = true; })
      .def("set_insertion_point_after", [](Program& p, int op_id) {
        // ops added from now on are placed right after `op_id` in program order (in the order they are added)
        for (size_t i = 0; i < p.order.size(); ++i)
          if (p.order[i] == op_id) { p.insert_at = (int)i + 1; return; }
        throw std::runtime_error("ir: unknown op " + std::to_string(op_id));
      })
      .def("set_insertion_point_before", [](Program& p, int op_id) {
        for (size_t i = 0; i < p.order.size(); ++i)
          if (p.order[i] == op_id) { p.insert_at = (int)i; return; }
        throw std::runtime_error("ir: unknown op " + std::to_string(op_id));
      })
      .def("reset_insertion_point", [](Program& p) { p.insert_at = -1; })
      .def("clone", [](const Program& p) {
        Out os;
        print_program(p, os, 0);
        const std::string text = os.str();
        Parser ps(text);
        return ps.program();
      })
      .def("memory_plan", [](const Program& p, int64_t align) {
        const MemoryPlan plan = plan_memory(p, align);
        py::dict d, off;
        for (const auto& kv : plan.offset) off[py::int_(kv.first)] = kv.second;
        d["offsets"] = off; d["peak_bytes"] = plan.peak; d["naive_bytes"] = plan.naive;
        return d;
      }, py::arg("align") = 256)
      .def("__str__", [](const Program& p) { Out os; print_program(p, os, 0); return os.str(); })
      .def_static("parse", [](const std::string& text) { Parser ps(text); return ps.program(); });

  py::class_<PassManager>(m, "IrPassManager")
      .def(py::init<std::vector<std::string>>(), py::arg("passes") = std::vector<std::string>())
      .def("add_pass", &PassManager::add_pass)
      .def("add_pattern", &PassManager::add_pattern)
      .def("set_folder", &PassManager::set_folder)
      .def("set_type_infer", &PassManager::set_type_infer)
      .def("enable_ir_printing", &PassManager::enable_ir_printing, py::arg("on") = true)
      .def("run", &PassManager::run);
}

}  // namespace runtime
}  // namespace b200
