"""paddle_b200.pir: native SSA IR, printer / parser round trip, DCE / CSE / identity / constant folding / inplace passes, declarative rewrite
patterns, the memory plan, and the static.Program -> IR -> static.Program path (results identical, fewer nodes).
Parity model: test/ir/pir/*, test/cpp/pir/pattern_rewrite/*."""
import numpy as np
import pytest
import torch

import paddle_b200 as paddle
from paddle_b200 import pir

pytestmark = pytest.mark.skipif(not pir.core_available(), reason="native extension not built")


def _mlp_ir():
    p = pir.Program()
    x = p.add_input("x", "float32", [4, 8])
    w = p.add_param("w", "float32", [8, 16])
    b = p.add_param("b", "float32", [16])
    (t,) = p.add_op("pd_op.matmul", [x, w], {}, [("float32", [4, 16])])
    (y,) = p.add_op("pd_op.add", [t, b], {}, [("float32", [4, 16])])
    (z,) = p.add_op("pd_op.gelu", [y], {}, [("float32", [4, 16])])
    (dead,) = p.add_op("pd_op.exp", [y], {}, [("float32", [4, 16])])
    (d2,) = p.add_op("pd_op.neg", [dead], {}, [("float32", [4, 16])])
    p.set_outputs([z])
    return p, (x, w, b, t, y, z)


def test_build_print_parse_round_trip():
    p, _ = _mlp_ir()
    p.verify()
    text = str(p)
    assert "pd_op.matmul" in text and "tensor<4x16xfloat32>" in text and "return" in text
    q = pir.parse(text)
    assert str(q) == str(pir.parse(str(q)))
    assert q.num_ops() == p.num_ops()
    assert [o["name"] for o in q.ops()] == [o["name"] for o in p.ops()]
    # attributes of every kind survive the text form
    r = pir.Program()
    a = r.add_input("a", "bfloat16", [2, -1, 3])
    r.add_op("pd_op.foo", [a], {"i": 3, "f": 0.5, "neg": -2, "s": 'he"llo', "flag": True, "li": [1, 2, 3], "lf": [0.5, 1.5]}, [("bfloat16", [2, -1, 3])])
    r.set_outputs([r.ops()[0]["results"][0]])
    r2 = pir.parse(str(r))
    assert r2.ops()[0]["attrs"] == r.ops()[0]["attrs"]
    assert r2.value_type(r2.ops()[0]["results"][0]) == ("bfloat16", [2, -1, 3])


def test_verify_rejects_use_before_def_and_unknown_values():
    p = pir.Program()
    x = p.add_input("x", "float32", [2])
    with pytest.raises(RuntimeError):
        p.add_op("pd_op.relu", [x + 7], {}, [("float32", [2])])
    with pytest.raises(RuntimeError):
        pir.parse('program {\n %0 = input "x" : tensor<2xfloat32>\n %1 = pd_op.relu(%5) : tensor<2xfloat32>\n return %1\n}')


def test_dce_cse_identity_passes():
    p, (x, w, b, t, y, z) = _mlp_ir()
    # duplicate subexpression + no-op reshape + transpose pair
    (t2,) = p.add_op("pd_op.matmul", [x, w], {}, [("float32", [4, 16])])
    (y2,) = p.add_op("pd_op.add", [t2, b], {}, [("float32", [4, 16])])
    (r,) = p.add_op("pd_op.reshape", [y2], {"a1": [4, 16]}, [("float32", [4, 16])])
    (tr1,) = p.add_op("pd_op.transpose", [r], {"perm": [1, 0]}, [("float32", [16, 4])])
    (tr2,) = p.add_op("pd_op.transpose", [tr1], {"perm": [1, 0]}, [("float32", [4, 16])])
    (s,) = p.add_op("pd_op.sub", [tr2, z], {}, [("float32", [4, 16])])
    p.set_outputs([s])
    pm = pir.PassManager(["identity_elim", "cse", "dce", "compact"], patterns=[])
    rep = pm.run(p)
    assert {r_["pass"] for r_ in rep} == {"identity_elim", "cse", "dce", "compact"}
    names = [o["name"] for o in p.ops()]
    assert names.count("pd_op.matmul") == 1 and names.count("pd_op.add") == 1
    assert "pd_op.reshape" not in names and "pd_op.transpose" not in names and "pd_op.exp" not in names and "pd_op.neg" not in names
    assert names == ["pd_op.matmul", "pd_op.add", "pd_op.gelu", "pd_op.sub"]
    p.verify()


def test_impure_ops_survive_dce_and_cse():
    p = pir.Program()
    x = p.add_input("x", "float32", [4])
    (a,) = p.add_op("pd_op.dropout", [x], {"p": 0.5}, [("float32", [4])])
    (b,) = p.add_op("pd_op.dropout", [x], {"p": 0.5}, [("float32", [4])])
    p.add_op("pd_op.add_", [x, a], {}, [("float32", [4])])       # in-place: side effect, result unused
    (c,) = p.add_op("pd_op.add", [a, b], {}, [("float32", [4])])
    p.set_outputs([c])
    pir.PassManager(["cse", "dce"], patterns=[]).run(p)
    names = [o["name"] for o in p.ops()]
    assert names.count("pd_op.dropout") == 2 and "pd_op.add_" in names


def test_rewrite_patterns_fuse_linear_and_swiglu():
    p, (x, w, b, t, y, z) = _mlp_ir()
    g = p.add_param("g", "float32", [4, 16])
    (s,) = p.add_op("pd_op.silu", [z], {}, [("float32", [4, 16])])
    (m,) = p.add_op("pd_op.mul", [s, g], {}, [("float32", [4, 16])])
    p.set_outputs([m])
    rep = pir.PassManager().run(p)
    names = [o["name"] for o in p.ops()]
    assert names == ["fused_linear", "swiglu"], names
    fl = p.ops()[0]
    assert fl["operands"] == [x, w, b] and fl["attrs"].get("activation") == "gelu"
    assert any(r["pass"] == "fuse_matmul_add" and r["changed"] == 1 for r in rep)
    p.verify()


def test_pattern_needs_single_use_of_inner_values():
    p, (x, w, b, t, y, z) = _mlp_ir()
    (u,) = p.add_op("pd_op.relu", [t], {}, [("float32", [4, 16])])      # the matmul result has a second consumer
    (o,) = p.add_op("pd_op.add", [u, z], {}, [("float32", [4, 16])])
    p.set_outputs([o])
    pir.PassManager(["fuse_matmul_add", "dce"]).run(p)
    assert "pd_op.matmul" in [o_["name"] for o_ in p.ops()] and "fused_linear" not in [o_["name"] for o_ in p.ops()]


def test_custom_pattern_with_attribute_constraint_and_copy():
    p = pir.Program()
    x = p.add_input("x", "float32", [4, 8])
    (a,) = p.add_op("pd_op.scale", [x], {"scale": 2.0}, [("float32", [4, 8])])
    (b,) = p.add_op("pd_op.softmax", [a], {"axis": -1}, [("float32", [4, 8])])
    (c,) = p.add_op("pd_op.scale", [x], {"scale": 3.0}, [("float32", [4, 8])])
    (d,) = p.add_op("pd_op.softmax", [c], {"axis": 0}, [("float32", [4, 8])])
    (e,) = p.add_op("pd_op.add", [b, d], {}, [("float32", [4, 8])])
    p.set_outputs([e])
    pm = pir.PassManager([], patterns=[])
    pm.add_pattern("fuse_scale2_softmax", [("scale", ["x"], ["t"], {"scale": 2.0}), ("softmax", ["t"], ["y"])],
                   [("scaled_softmax", ["x"], ["y"], {"factor": 2.0, "axis": "$y.axis"})])
    pm.run(p)
    ops = {o["name"]: o for o in p.ops()}
    assert "scaled_softmax" in ops and ops["scaled_softmax"]["attrs"] == {"factor": 2.0, "axis": -1}
    assert [o["name"] for o in p.ops()].count("pd_op.softmax") == 1       # the scale = 3.0 branch does not match


def test_constant_folding_through_python_callback():
    p = pir.Program()
    pm = pir.PassManager(["constant_fold", "dce"], patterns=[])
    c1, c2 = pm.add_constant(torch.tensor([1.0, 2.0])), pm.add_constant(torch.tensor([3.0, 5.0]))
    x = p.add_input("x", "float32", [2])
    (a,) = p.add_op("pd_op.constant", [], {"const_id": c1}, [("float32", [2])])
    (b,) = p.add_op("pd_op.constant", [], {"const_id": c2}, [("float32", [2])])
    (s,) = p.add_op("pd_op.add", [a, b], {}, [("float32", [2])])
    (e,) = p.add_op("pd_op.exp", [s], {}, [("float32", [2])])
    (y,) = p.add_op("pd_op.mul", [x, e], {}, [("float32", [2])])
    p.set_outputs([y])
    pm.run(p)
    ops = p.ops()
    assert [o["name"] for o in ops] == ["pd_op.constant", "pd_op.mul"]
    assert torch.allclose(pm.constant(ops[0]["attrs"]["const_id"]), torch.exp(torch.tensor([4.0, 7.0])))


def test_inplace_marking_and_memory_plan():
    p = pir.Program()
    x = p.add_input("x", "float32", [1024, 1024])
    (a,) = p.add_op("pd_op.exp", [x], {}, [("float32", [1024, 1024])])          # operand is a program input: never in place
    (b,) = p.add_op("pd_op.relu", [a], {}, [("float32", [1024, 1024])])         # a dies here
    (c,) = p.add_op("pd_op.tanh", [b], {}, [("float32", [1024, 1024])])
    (d,) = p.add_op("pd_op.add", [c, b], {}, [("float32", [1024, 1024])])       # b still alive at tanh -> tanh not in place; c dies at add
    p.set_outputs([d])
    pir.PassManager(["inplace"], patterns=[]).run(p)
    marks = {o["name"]: o["attrs"].get("inplace", False) for o in p.ops()}
    assert marks == {"pd_op.exp": False, "pd_op.relu": True, "pd_op.tanh": False, "pd_op.add": True}
    plan = p.memory_plan(256)
    assert plan["naive_bytes"] == 4 * 4 * 1024 * 1024
    assert plan["peak_bytes"] <= 3 * 4 * 1024 * 1024          # a's slot is reused once a is dead
    offs = plan["offsets"]
    # simultaneously live values never overlap: b, c are both alive when d is produced
    assert len({offs[b], offs[c], offs[d]}) == 3


def test_regions_print_parse_and_verify():
    body = pir.Program()
    i = body.add_input("i", "int64", [])
    (j,) = body.add_op("pd_op.add", [i, i], {}, [("int64", [])])
    body.set_outputs([j])
    p = pir.Program()
    x = p.add_input("x", "int64", [])
    (y,) = p.add_op("pd_op.while", [x], {"max_iter": 4}, [("int64", [])])
    p.add_region(p.ops()[0]["id"], body)
    p.set_outputs([y])
    p.verify()
    q = pir.parse(str(p))
    assert q.ops()[0]["num_regions"] == 1 and [o["name"] for o in q.region(q.ops()[0]["id"], 0).ops()] == ["pd_op.add"]
    pir.PassManager(["dce", "cse"], patterns=[]).run(q)        # control flow is never removed
    assert q.num_ops() == 1


def test_static_program_through_pir_same_results_fewer_nodes():
    paddle.enable_static()
    try:
        main = paddle.static.Program()
        with paddle.static.program_guard(main):
            x = paddle.static.data("x", [4, 8], "float32")
            w = paddle.to_tensor(np.random.RandomState(0).randn(8, 16).astype("float32"))
            b = paddle.to_tensor(np.random.RandomState(1).randn(16).astype("float32"))
            g = paddle.to_tensor(np.random.RandomState(2).randn(4, 16).astype("float32"))
            h1 = paddle.add(paddle.matmul(x, w), b)
            h2 = paddle.add(paddle.matmul(x, w), b)                        # common subexpression
            unused = paddle.exp(h2)                                        # dead
            y = paddle.multiply(paddle.nn.functional.silu(h1), g) + paddle.reshape(h2, [4, 16])
        exe = paddle.static.Executor()
        feed = {"x": np.random.RandomState(3).randn(4, 8).astype("float32")}
        (ref,) = exe.run(main, feed=feed, fetch_list=[y])
        opt, report = pir.optimize(main, fetch_list=[y], return_report=True)
        assert len(opt.nodes) < len(main.nodes)
        names = [getattr(n.fn, "__name__", "") for n in opt.nodes]
        assert "_fused_linear" in names and "_swiglu" in names
        (got,) = exe.run(opt, feed=feed, fetch_list=[y])
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
        assert any(r["pass"] == "cse" and r["changed"] >= 1 for r in report)
    finally:
        paddle.disable_static()


def test_executor_runs_through_pir_when_flag_is_set():
    paddle.enable_static()
    try:
        main = paddle.static.Program()
        with paddle.static.program_guard(main):
            x = paddle.static.data("x", [3, 5], "float32")
            w = paddle.to_tensor(np.random.RandomState(0).randn(5, 7).astype("float32"))
            b = paddle.to_tensor(np.zeros(7, "float32") + 0.5)
            h = paddle.add(paddle.matmul(x, w), b)
            dup = paddle.add(paddle.matmul(x, w), b)
            y = paddle.nn.functional.relu(h) + dup
        exe = paddle.static.Executor()
        feed = {"x": np.random.RandomState(1).randn(3, 5).astype("float32")}
        (ref,) = exe.run(main, feed=feed, fetch_list=[y])
        paddle.set_flags({"FLAGS_enable_pir_api": True})
        try:
            (got,) = exe.run(main, feed=feed, fetch_list=[y])
            cached = list(main.__dict__["_pir_cache"].values())[0][1]
            assert len(cached.nodes) < len(main.nodes)
            (again,) = exe.run(main, feed=feed, fetch_list=[y])
        finally:
            paddle.set_flags({"FLAGS_enable_pir_api": False})
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(again, ref, rtol=1e-5, atol=1e-5)
    finally:
        paddle.disable_static()


def test_inference_predictor_runs_ir_passes_on_program_artifacts(tmp_path):
    """Config.pass_builder() names map onto the native IR passes and are really applied to a saved program (fewer ops, same output);
    switch_ir_optim(False) leaves the program as saved."""
    from paddle_b200 import inference, static

    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [2, 4], "float32")
            w = paddle.to_tensor(np.random.RandomState(0).randn(4, 6).astype("float32"))
            b = paddle.to_tensor(np.random.RandomState(1).randn(6).astype("float32"))
            h = paddle.add(paddle.matmul(x, w), b)
            h2 = paddle.add(paddle.matmul(x, w), b)
            y = paddle.nn.functional.gelu(h) + paddle.reshape(h2, [2, 6])
        exe = static.Executor()
        static.save_inference_model(str(tmp_path / "m"), [x], [y], exe, program=main)
    finally:
        paddle.disable_static()
    data = np.random.RandomState(2).randn(2, 4).astype("float32")
    cfg0 = inference.Config(str(tmp_path / "m.pdmodel"), str(tmp_path / "m.pdiparams"))
    cfg0.switch_ir_optim(False)
    p0 = inference.create_predictor(cfg0)
    assert p0.ir_pass_report() == []
    ref = p0.run([data])[0]
    n_saved = len(p0._layer._blob["program"].nodes)
    cfg = inference.Config(str(tmp_path / "m.pdmodel"), str(tmp_path / "m.pdiparams"))
    assert "fuse_gemm_epilogue_pass" in cfg.pass_builder().all_passes()
    cfg.pass_builder().delete_pass("inplace_pass")
    p1 = inference.create_predictor(cfg)
    rep = p1.ir_pass_report()
    assert rep and {r["pass"] for r in rep} >= {"cse", "dce", "fuse_matmul_add"} and "inplace" not in {r["pass"] for r in rep}
    assert len(p1._layer._blob["program"].nodes) < n_saved
    got = p1.run([data])[0]
    np.testing.assert_allclose(np.asarray(got), np.asarray(ref), rtol=1e-5, atol=1e-5)


def test_programs_with_run_time_control_flow_go_through_the_passes():
    """A while_loop node is opaque to the IR (side-effect op that keeps its inputs alive); the ops around it are still optimised, and when CSE
    moves one of the loop's inputs to another slot the lowering binds the old slot to the survivor."""
    paddle.enable_static()
    try:
        main = paddle.static.Program()
        with paddle.static.program_guard(main):
            x = paddle.static.data("x", [4], "float32")
            n = paddle.static.data("n", [1], "int64")
            a = paddle.exp(x)
            b = paddle.exp(x)                                   # same value as `a`: CSE merges it, the loop below reads `b`
            i0 = paddle.zeros([1], "int64")

            def cond(i, acc):
                return i < n

            def body(i, acc):
                return i + 1, acc + b

            _, acc = paddle.static.nn.while_loop(cond, body, [i0, paddle.zeros([4], "float32")])
            dead = paddle.tanh(a)
            y = acc * a
        exe = paddle.static.Executor()
        feed = {"x": np.array([0.1, 0.2, 0.3, 0.4], "float32"), "n": np.array([3], "int64")}
        (ref,) = exe.run(main, feed=feed, fetch_list=[y])
        np.testing.assert_allclose(ref, 3 * np.exp(feed["x"]) ** 2, rtol=1e-5)
        opt, report = pir.optimize(main, fetch_list=[y], return_report=True)
        assert any(r["pass"] == "cse" and r["changed"] >= 1 for r in report) and any(r["pass"] == "dce" and r["changed"] >= 1 for r in report)
        assert sum(1 for nd in opt.nodes if getattr(nd.fn, "__name__", "") == "exp") == 1
        assert any(nd.kind == "control" for nd in opt.nodes)
        (got,) = exe.run(opt, feed=feed, fetch_list=[y])
        np.testing.assert_allclose(got, ref, rtol=1e-6)
        feed2 = {"x": feed["x"], "n": np.array([5], "int64")}      # the trip count is still a run-time value
        (got2,) = exe.run(opt, feed=feed2, fetch_list=[y])
        np.testing.assert_allclose(got2, 5 * np.exp(feed["x"]) ** 2, rtol=1e-5)
    finally:
        paddle.disable_static()


def test_object_views_walk_and_edit_a_program():
    p, (x, w, b, t, y, z) = _mlp_ir()
    blk = pir.global_block(p)
    assert len(blk) == 5 and [o.name() for o in blk][:3] == ["pd_op.matmul", "pd_op.add", "pd_op.gelu"]
    assert [v.name for v in blk.args()] == ["x", "w", "b"] and set(blk.kwargs()) == {"x", "w", "b"}
    mm = blk.ops[0]
    assert mm.num_operands() == 2 and mm.num_results() == 1 and mm.operand_source(0) == pir.Value(p, x)
    out = mm.result(0)
    assert out.shape == [4, 16] and out.dtype == "float32" and not out.is_block_argument() and pir.Value(p, x).is_block_argument()
    assert out.get_defining_op().name() == "pd_op.matmul" and [o.name() for o in out.all_used_ops()] == ["pd_op.add"]
    yv = pir.Value(p, y)
    assert yv.use_count() == 2 and {o.name() for o in yv.all_used_ops()} == {"pd_op.gelu", "pd_op.exp"}
    # manual rewrite through the views: gelu reads the matmul result directly, the add becomes dead
    yv.replace_all_uses_with(out)
    assert yv.use_empty()
    blk.ops[1].erase()
    pir.PassManager(["dce", "compact"], patterns=[]).run(p)
    assert [o.name() for o in pir.global_block(p)] == ["pd_op.matmul", "pd_op.gelu"]
    assert "Value(%" in repr(out) and "Operation(pd_op.matmul" in repr(pir.global_block(p).ops[0])


def test_conv_bn_fuse_in_the_predictor(tmp_path):
    """The predictor's default passes fold inference batch norms into the preceding convolution (fewer ops, same numbers)."""
    from paddle_b200 import inference, static

    paddle.seed(3)
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [2, 3, 10, 10], "float32")
            c1, b1 = paddle.nn.Conv2D(3, 8, 3, padding=1), paddle.nn.BatchNorm2D(8)
            c2, b2 = paddle.nn.Conv2D(8, 8, 3, padding=1, bias_attr=False), paddle.nn.BatchNorm2D(8)
            rs = np.random.RandomState(5)
            for bn in (b1, b2):
                bn._mean.set_value(rs.randn(8).astype("float32") * 0.2)          # concrete values: set_value is an assignment, not an op of the program
                bn._variance.set_value(rs.rand(8).astype("float32") + 0.5)
                bn.weight.set_value(rs.randn(8).astype("float32"))
                bn.bias.set_value(rs.randn(8).astype("float32"))
                bn.eval()
            h = paddle.nn.functional.relu(b1(c1(x)))
            y = paddle.nn.functional.relu(b2(c2(h)) + h)
        static.save_inference_model(str(tmp_path / "m"), [x], [y], static.Executor(), program=main)
    finally:
        paddle.disable_static()
    data = np.random.RandomState(0).randn(2, 3, 10, 10).astype("float32")
    cfg0 = inference.Config(str(tmp_path / "m.pdmodel"), str(tmp_path / "m.pdiparams"))
    cfg0.switch_ir_optim(False)
    ref = inference.create_predictor(cfg0).run([data])[0]
    cfg = inference.Config(str(tmp_path / "m.pdmodel"), str(tmp_path / "m.pdiparams"))
    assert cfg.pass_builder().all_passes()[0] == "conv_bn_fuse_pass"
    p = inference.create_predictor(cfg)
    rep = {r["pass"]: r for r in p.ir_pass_report()}
    assert rep["conv_bn_fuse"]["changed"] == 2
    names = [getattr(n.fn, "__name__", "") for n in p._layer._blob["program"].nodes]
    assert "batch_norm" not in names and names.count("conv2d") == 2
    np.testing.assert_allclose(np.asarray(p.run([data])[0]), np.asarray(ref), rtol=1e-4, atol=1e-5)
