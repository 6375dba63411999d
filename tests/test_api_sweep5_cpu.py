"""Sixth sweep: static-graph leftovers (serialization, scopes, EMA, control flow), jit / quantization / profiler / utils /
distributed single-process names."""
import os

import numpy as np
import pytest

import paddle_b200 as paddle

rng = np.random.RandomState(51)


def t(a):
    return paddle.to_tensor(np.asarray(a))


@pytest.fixture
def static_mode():
    paddle.enable_static()
    try:
        yield
    finally:
        paddle.disable_static()


def _tiny_program():
    S = paddle.static
    main, start = S.Program(), S.Program()
    with S.program_guard(main, start):
        x = S.data("x", [-1, 4], "float32")
        with S.name_scope("block"):
            h = S.nn.fc(x, 3, activation="relu")
        y = paddle.sum(h, axis=1)
    return main, start, x, y


def test_static_serialization_and_state(static_mode, tmp_path):
    S = paddle.static
    main, start, x, y = _tiny_program()
    exe = S.Executor(S.cpu_places(1)[0])
    exe.run(start)
    feed = {"x": rng.randn(2, 4).astype("float32")}
    ref, = exe.run(main, feed=feed, fetch_list=[y])
    assert S.default_main_program() is not None and S.default_startup_program() is not None and S.in_static_mode() and not S.in_dynamic_mode()
    prog_bytes = S.serialize_program([x], [y], program=main)
    pers_bytes = S.serialize_persistables([x], [y], exe, program=main)
    S.save_to_file(str(tmp_path / "m.pdmodel"), prog_bytes)
    S.save_to_file(str(tmp_path / "m.pdiparams"), pers_bytes)
    prog2 = S.deserialize_program(S.load_from_file(str(tmp_path / "m.pdmodel")))
    S.deserialize_persistables(prog2, S.load_from_file(str(tmp_path / "m.pdiparams")), exe)
    out_d, = exe.run(prog2, feed=feed, fetch_list=prog2._fetch_vars)
    np.testing.assert_allclose(out_d, ref, rtol=1e-6)
    assert prog2._feed_names == ["x"]
    norm = S.normalize_program(main, [x], [y])
    assert norm is not None
    S.save(main, str(tmp_path / "ckpt"))
    state = S.load_program_state(str(tmp_path / "ckpt"))
    assert state and all(isinstance(v, np.ndarray) for v in state.values())
    zeroed = {k: np.zeros_like(v) for k, v in state.items()}
    S.set_program_state(main, zeroed)
    out0, = exe.run(main, feed=feed, fetch_list=[y])
    assert np.allclose(out0, 0)
    S.set_program_state(main, state)
    out1, = exe.run(main, feed=feed, fetch_list=[y])
    np.testing.assert_allclose(out1, ref, rtol=1e-6)
    cp = S.CompiledProgram(main, build_strategy=S.BuildStrategy())
    out2, = exe.run(cp, feed=feed, fetch_list=[y])
    np.testing.assert_allclose(out2, ref, rtol=1e-6)
    assert S.ExecutionStrategy() is not None and isinstance(S.cuda_places(), list)
    with S.scope_guard(S.global_scope()):
        with S.device_guard("cpu"):
            pass
    g = S.create_global_var([2], 1.5, "float32", persistable=True, name="gv")
    assert np.allclose(np.asarray(g), 1.5)
    assert isinstance(x, S.Variable)
    assert S.WeightNormParamAttr(dim=0) is not None


def test_static_backward_ema_print_control_flow(static_mode):
    S = paddle.static
    main, start = S.Program(), S.Program()
    with S.program_guard(main, start):
        x = S.data("x", [3, 2], "float32")
        w = S.create_parameter([2, 1], "float32", name="w_cf")
        loss = paddle.mean(paddle.matmul(x, w))
        S.Print(loss, message="loss")
        pg = S.append_backward(loss)
        assert pg and pg[0][0] is not None
        i = paddle.full([1], 0, "int64")
        ten = paddle.full([1], 5, "int64")
        (i_out,) = S.nn.while_loop(lambda i: i < ten, lambda i: [i + 1], [i])
        sw = S.nn.switch_case(paddle.full([1], 1, "int32"), {0: lambda: paddle.full([1], 10.0), 1: lambda: paddle.full([1], 20.0)}, default=lambda: paddle.full([1], -1.0))
        pf = S.nn.py_func(lambda a: a * 2, x, None)
        sp = S.nn.static_pylayer(lambda a: a + 1, [x])
    exe = S.Executor()
    exe.run(start)
    outs = exe.run(main, feed={"x": np.ones((3, 2), "float32")}, fetch_list=[i_out, sw, pf, sp])
    assert int(outs[0][0]) == 5 and float(outs[1][0]) == 20.0 and np.allclose(outs[2], 2) and np.allclose(outs[3], 2)
    ema = S.ExponentialMovingAverage(0.5)
    lin_w = paddle.create_parameter([2], "float32")
    ema.update([lin_w]) if "parameters" in ema.update.__code__.co_varnames or ema.update.__code__.co_argcount > 1 else ema.update()
    with ema.apply(exe):
        pass
    ema.restore(exe)


def test_static_nn_leftovers(static_mode):
    S = paddle.static
    main, start = S.Program(), S.Program()
    with S.program_guard(main, start):
        a = S.data("a", [4, 3], "float32")
        b = S.data("b", [4, 5], "float32")
        btp = S.nn.bilinear_tensor_product(a, b, 6)
        dn = S.nn.data_norm(a)
        rc = S.nn.row_conv(S.data("seq", [2, 7, 3], "float32"), 2)
        emb = S.nn.sparse_embedding(S.data("ids", [4, 1], "int64"), [10, 8])
        lbl = S.data("lbl", [4, 1], "int64")
        nce = S.nn.nce(a, lbl, 20, num_neg_samples=5)
    exe = S.Executor()
    exe.run(start)
    outs = exe.run(main, feed={"a": rng.randn(4, 3).astype("float32"), "b": rng.randn(4, 5).astype("float32"), "seq": rng.randn(2, 7, 3).astype("float32"),
                               "ids": rng.randint(0, 10, (4, 1)), "lbl": rng.randint(0, 20, (4, 1))}, fetch_list=[btp, dn, rc, emb, nce])
    assert outs[0].shape == (4, 6) and outs[1].shape == (4, 3) and outs[2].shape == (2, 7, 3) and outs[3].shape[-1] == 8 and outs[4].shape[0] == 4
    assert all(np.isfinite(o).all() for o in outs)


def test_jit_leftovers(tmp_path):
    J = paddle.jit
    net = paddle.nn.Linear(4, 2)
    J.save(net, str(tmp_path / "n"), input_spec=[paddle.static.InputSpec([None, 4], "float32")])
    loaded = J.load(str(tmp_path / "n"))
    assert isinstance(loaded, J.TranslatedLayer) or callable(loaded)
    x = t(rng.randn(3, 4).astype("float32"))
    np.testing.assert_allclose(loaded(x).numpy(), net(x).numpy(), rtol=1e-6)
    J.set_code_level(50)
    J.set_verbosity(0)
    import math

    J.ignore_module([math])

    @J.not_to_static
    def helper(v):
        return v * 2

    @J.to_static
    def fn(v):
        return helper(v) + 1

    np.testing.assert_allclose(fn(x).numpy(), x.numpy() * 2 + 1, rtol=1e-6)
    J.enable_to_static(False)
    try:
        np.testing.assert_allclose(fn(x).numpy(), x.numpy() * 2 + 1, rtol=1e-6)
    finally:
        J.enable_to_static(True)


def test_quantization_leftovers():
    Q = paddle.quantization
    from paddle_b200.quantization import quanter  # noqa: F401

    class MyObserver(Q.BaseObserver):
        def __init__(self):
            super().__init__()
            self.m = 0.0

        def forward(self, x):
            self.m = max(self.m, float(x.abs().max()))
            return x

        def cal_thresholds(self):
            pass

        def scales(self):
            return paddle.to_tensor(self.m / 127.0)

        def zero_points(self):
            return paddle.to_tensor(0)

        def bit_length(self):
            return 8

        def quant_axis(self):
            return -1

    ob = MyObserver()
    ob(t(np.array([1.0, -3.0], "float32")))
    assert abs(float(ob.scales()) - 3 / 127) < 1e-6 and isinstance(ob, Q.BaseQuanter)
    assert paddle.nn.quant.Stub() is not None


def test_profiler_leftovers(tmp_path):
    P = paddle.profiler
    assert P.ProfilerState.RECORD is not None and P.SortedKeys.GPUTotal is not None and P.SummaryView.KernelView is not None and P.TracerEventType.Operator is not None
    prof = P.Profiler(targets=[P.ProfilerTarget.CPU], scheduler=(1, 3), on_trace_ready=P.export_protobuf(str(tmp_path / "pb")), timer_only=False)
    prof.start()
    for _ in range(4):
        with P.profile_range("step"):
            paddle.matmul(t(rng.randn(8, 8)), t(rng.randn(8, 8)))
        prof.step(num_samples=8)
    prof.stop()
    assert "ips" in prof.step_info(unit="samples")
    out = str(tmp_path / "t.json")
    prof.export(out)
    assert "traceEvents" in P.load_profiler_result(out)
    prof.summary(sorted_by=P.SortedKeys.CPUTotal, views=[P.SummaryView.OperatorView])


def test_utils_leftovers(tmp_path):
    U = paddle.utils
    assert U.try_import("math") is not None
    with pytest.raises(ImportError):
        U.try_import("surely_not_a_module_xyz")
    U.require_version("0.0.1")
    with pytest.raises(Exception):
        U.require_version("999.0.0")

    @U.deprecated(since="2.0", update_to="paddle.new_fn", reason="test")
    def old(a):
        return a + 1

    with pytest.warns(Warning):
        assert old(1) == 2
    with U.unique_name.guard():
        a, b = U.unique_name.generate("fc"), U.unique_name.generate("fc")
    assert a != b and a.startswith("fc")
    U.run_check()
    assert hasattr(U.cpp_extension, "load") and hasattr(U.cpp_extension, "CUDAExtension") and hasattr(U.download, "get_weights_path_from_url")


def test_distributed_single_process_names(tmp_path):
    D = paddle.distributed
    assert D.ReduceOp.SUM is not None and D.ParallelEnv().world_size == 1 and D.get_world_size() == 1 and D.get_rank() == 0
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ["MASTER_PORT"] = str(29000 + os.getpid() % 2000)
    D.init_parallel_env()
    try:
        g = D.new_group([0])
        assert isinstance(g, D.Group) and D.get_group(g.id) is g and D.get_backend() in ("gloo", "nccl")
        x = t(np.array([1.0, 2.0], "float32"))
        D.all_reduce(x)
        out = []
        D.all_gather(out, x)
        assert len(out) == 1
        big = paddle.zeros([2], "float32")
        D.all_gather_into_tensor(big, x)
        np.testing.assert_allclose(big.numpy(), x.numpy())
        rs = paddle.zeros([2], "float32")
        D.reduce_scatter(rs, [x])
        D.reduce_scatter_tensor(rs, x)
        a2a = paddle.zeros([2], "float32")
        D.alltoall_single(x, a2a)
        D.stream.alltoall_single(a2a, x, use_calc_stream=True)
        with pytest.raises(RuntimeError):
            D.stream.all_reduce(x, sync_op=False, use_calc_stream=True)
        np.testing.assert_allclose(a2a.numpy(), x.numpy())
        objs = [{"a": 1}]
        D.broadcast_object_list(objs, 0)
        outl = [None]
        D.scatter_object_list(outl, [{"b": 2}], 0)
        assert objs[0] == {"a": 1} and outl[0] == {"b": 2}
        assert D.batch_isend_irecv is not None and D.P2POp is not None and D.isend is not None and D.irecv is not None
    finally:
        D.destroy_process_group()


def test_executor_train_from_dataset(static_mode, tmp_path):
    """Executor.train_from_dataset over an InMemoryDataset (multi-slot text) trains a static program."""
    from paddle_b200 import _build
    from paddle_b200.distributed import InMemoryDataset

    if _build.load(required=False) is None:
        pytest.skip("native extension not built")
    S = paddle.static
    w_true = np.array([0.5, -1.0, 2.0], "float32")
    lines = []
    for _ in range(64):
        x = rng.randn(3).astype("float32")
        lines.append(f"3 {x[0]} {x[1]} {x[2]} 1 {float(x @ w_true)}\n")
    p = tmp_path / "part-0"
    p.write_text("".join(lines))
    main, start = S.Program(), S.Program()
    with S.program_guard(main, start):
        feat = S.data("feat", [-1, 3], "float32")
        label = S.data("label", [-1, 1], "float32")
        pred = S.nn.fc(feat, 1)
        loss = paddle.mean((pred - label) ** 2)
        paddle.optimizer.SGD(0.1).minimize(loss)
    exe = S.Executor()
    exe.run(start)
    ds = InMemoryDataset()
    ds.init(batch_size=16, use_var=[feat, label])
    ds.set_filelist([str(p)])
    ds.load_into_memory()
    first = exe.train_from_dataset(main, ds, fetch_list=[loss])
    for _ in range(10):
        last = exe.train_from_dataset(main, ds, fetch_list=[loss])
    assert float(last[-1][0]) < float(first[0][0]) * 0.1
    inf = exe.infer_from_dataset(main, ds, fetch_list=[pred])
    assert len(inf) == 4 and inf[0][0].shape == (16, 1)


def test_static_amp_decorate(static_mode):
    S = paddle.static
    import paddle_b200.static.amp as samp

    main, start = S.Program(), S.Program()
    with S.program_guard(main, start):
        x = S.data("x", [8, 4], "float32")
        y = S.data("y", [8, 1], "float32")
        pred = S.nn.fc(x, 1)
        loss = paddle.mean((pred - y) ** 2)
        opt = samp.decorate(paddle.optimizer.SGD(0.1), samp.AutoMixedPrecisionLists(custom_black_list=["mean"]), dtype="bfloat16", level="O1")
        opt.minimize(loss)
    assert main._dist_attrs["amp"]["dtype"] == "bfloat16" and main._dist_attrs["amp"]["wrapped"] >= 1
    exe = S.Executor()
    exe.run(start)
    xs, ws = rng.randn(8, 4).astype("float32"), np.array([[1.0], [-2.0], [0.5], [0.0]], "float32")
    first = exe.run(main, feed={"x": xs, "y": xs @ ws}, fetch_list=[loss])[0]
    for _ in range(40):
        last = exe.run(main, feed={"x": xs, "y": xs @ ws}, fetch_list=[loss])[0]
    assert float(last) < float(first) * 0.3
    with pytest.raises(ValueError):
        samp.AutoMixedPrecisionLists(custom_white_list=["matmul"], custom_black_list=["matmul"])
    assert samp.bf16.AutoMixedPrecisionListsBF16 is samp.AutoMixedPrecisionLists
