"""Static-graph mode: Program recording, Executor replay, training through minimize, passes, inference model IO.
Parity: test/legacy_test/test_executor_*.py, test_program.py, test/ir/pir/*pass*."""
import numpy as np
import pytest

import paddle_b200 as paddle
from paddle_b200 import static


@pytest.fixture(autouse=True)
def _static_mode():
    paddle.enable_static()
    yield
    paddle.disable_static()


def test_program_record_and_run():
    main, start = static.Program(), static.Program()
    with static.program_guard(main, start):
        x = static.data("x", [-1, 4], "float32")
        y = paddle.tanh(x * 2.0 + 1.0).sum(axis=1)
    exe = static.Executor()
    exe.run(start)
    for bs in (2, 5):
        a = np.random.rand(bs, 4).astype("float32")
        out, = exe.run(main, feed={"x": a}, fetch_list=[y])
        np.testing.assert_allclose(out, np.tanh(a * 2 + 1).sum(1), rtol=1e-5)


def test_static_training_converges():
    paddle.seed(0)
    main, start = static.Program(), static.Program()
    with static.program_guard(main, start):
        x = static.data("x", [-1, 3], "float32")
        t = static.data("t", [-1, 1], "float32")
        h = static.nn.fc(x, 8, activation="tanh")
        pred = static.nn.fc(h, 1)
        loss = paddle.nn.functional.mse_loss(pred, t)
        paddle.optimizer.Adam(0.05).minimize(loss)
    exe = static.Executor()
    exe.run(start)
    rng = np.random.RandomState(0)
    X = rng.rand(64, 3).astype("float32")
    T = (X @ np.array([[1.0], [-2.0], [0.5]], dtype="float32"))
    losses = [float(exe.run(main, feed={"x": X, "t": T}, fetch_list=[loss])[0]) for _ in range(60)]
    assert losses[-1] < losses[0] * 0.2
    test_prog = main.clone(for_test=True)
    l1 = float(exe.run(test_prog, feed={"x": X, "t": T}, fetch_list=[loss])[0])
    l2 = float(exe.run(test_prog, feed={"x": X, "t": T}, fetch_list=[loss])[0])
    assert abs(l1 - l2) < 1e-7   # no update in the test clone


def test_passes():
    from paddle_b200.static import passes

    main = static.Program()
    w = paddle.to_tensor(np.random.rand(4, 6).astype("float32"))
    b = paddle.to_tensor(np.random.rand(6).astype("float32"))
    with static.program_guard(main):
        x = static.data("x", [-1, 4], "float32")
        c = paddle.ones([6]) * 3.0 + 1.0          # constant sub-graph
        y1 = paddle.matmul(x, w) + b               # gemm + bias
        e1 = paddle.exp(x)
        e2 = paddle.exp(x)                         # CSE
        dead = paddle.sin(x) * 2.0                 # unused
        out = y1 * c + (e1 + e2).sum(axis=1, keepdim=True)
    n0 = len(main.nodes)
    a = np.random.rand(3, 4).astype("float32")
    exe = static.Executor()
    ref, = exe.run(main, feed={"x": a}, fetch_list=[out])
    keep = {main._fetch_alias[id(out)]}
    pm = passes.PassManager()
    pm.apply(main, keep)
    assert pm.stats["dead_code_elimination"] >= 2 and pm.stats["common_subexpression_elimination"] >= 1
    # constants built from literals never enter the tape (they are evaluated while recording), so nothing is left to fold here
    assert pm.stats["constant_folding"] == 0 and pm.stats["fuse_gemm_epilogue"] == 1
    assert len(main.nodes) < n0
    got, = exe.run(main, feed={"x": a}, fetch_list=[out])
    np.testing.assert_allclose(got, ref, rtol=1e-5)
    np.testing.assert_allclose(ref, (a @ w.numpy() + b.numpy()) * 4.0 + 2 * np.exp(a).sum(1, keepdims=True), rtol=1e-5)


def test_save_load_inference_model(tmp_path):
    paddle.seed(1)
    main, start = static.Program(), static.Program()
    with static.program_guard(main, start):
        x = static.data("x", [-1, 5], "float32")
        y = static.nn.fc(x, 3, activation="relu")
    exe = static.Executor()
    exe.run(start)
    a = np.random.rand(2, 5).astype("float32")
    ref, = exe.run(main, feed={"x": a}, fetch_list=[y])
    static.save_inference_model(str(tmp_path / "m"), [x], [y], exe, program=main)
    prog, feeds, fetches = static.load_inference_model(str(tmp_path / "m"), exe)
    got, = exe.run(prog, feed={feeds[0]: a}, fetch_list=fetches)
    np.testing.assert_allclose(got, ref, rtol=1e-6)


def test_data_dependent_control_flow():
    """cond / case / switch_case / while_loop whose predicates are only known at run time (values come from feeds)."""
    import numpy as np

    import paddle_b200 as paddle

    paddle.enable_static()
    try:
        S = paddle.static
        main, start = S.Program(), S.Program()
        with S.program_guard(main, start):
            x = S.data("x", [1], "float32")
            n = S.data("n", [1], "int64")
            w = S.create_parameter([1], "float32", name="cf_w", default_initializer=paddle.nn.initializer.Constant(3.0))
            out = S.nn.cond(x > 0, lambda: x * w, lambda: x - 1)
            c = S.nn.case([(x > 5, lambda: x * 0 + 100), (x > 0, lambda: x * 0 + 10)], default=lambda: x * 0 - 1)
            sw = S.nn.switch_case(n, {0: lambda: x + 1, 1: lambda: x + 2}, default=lambda: x * 0)
            i0, acc0 = paddle.full([1], 0, "int64"), paddle.zeros([1])
            i_f, acc_f = S.nn.while_loop(lambda i, acc: i < n, lambda i, acc: [i + 1, acc + x], [i0, acc0])
            y = acc_f * 10
            loss = paddle.mean(out)
            grads = S.gradients([loss], [w])
        exe = S.Executor()
        exe.run(start)
        for v, k in ((3.0, 0), (-2.0, 1), (7.0, 4)):
            o, cc, s, i, acc, yy, g = exe.run(main, feed={"x": np.array([v], "float32"), "n": np.array([k])}, fetch_list=[out, c, sw, i_f, acc_f, y, grads[0]])
            assert float(o[0]) == (v * 3 if v > 0 else v - 1)
            assert float(cc[0]) == (100 if v > 5 else 10 if v > 0 else -1)
            assert float(s[0]) == {0: v + 1, 1: v + 2}.get(k, 0.0)
            assert int(i[0]) == k and float(acc[0]) == k * v and float(yy[0]) == 10 * k * v
            assert float(g[0]) == (v if v > 0 else 0.0)          # the gradient flows through the selected branch only
        # constant predicates keep the eager semantics (the untaken branch is not even traced)
        main2 = S.Program()
        with S.program_guard(main2):
            a = S.data("a", [1], "float32")
            r = S.nn.cond(paddle.full([1], 1.0) > 0, lambda: a + 1, lambda: 1 / 0)
        assert float(exe.run(main2, feed={"a": np.array([1.0], "float32")}, fetch_list=[r])[0][0]) == 2.0
    finally:
        paddle.disable_static()


def test_dynamic_batch_shapes_are_resolved_at_run_time():
    """x.shape reports -1 for dynamic dims; paddle.shape / numel and shape-taking ops (reshape, expand, full, arange, tile, slice) use the
    run-time extent, so one recorded program serves every batch size."""
    import numpy as np

    import paddle_b200 as paddle

    paddle.enable_static()
    try:
        S = paddle.static
        main = S.Program()
        with S.program_guard(main):
            x = S.data("x", [-1, 4], "float32")
            assert x.shape == [-1, 4]
            bsz = paddle.shape(x)[0]
            outs = [paddle.reshape(x, [bsz, 2, 2]), paddle.reshape(x, [x.shape[0], 2, 2]), paddle.expand(paddle.ones([1, 3]), [bsz, 3]), paddle.full([bsz, 2], 7.0),
                    paddle.arange(0, bsz), paddle.tile(paddle.ones([1, 2]), [bsz, 1]), paddle.slice(x, [0], [0], [bsz - 3]),
                    paddle.ones_like(x) * paddle.cast(paddle.numel(x), "float32"), S.nn.fc(x, 3)]
        exe = S.Executor()
        for b in (8, 5):
            res = exe.run(main, feed={"x": np.ones((b, 4), "float32")}, fetch_list=outs)
            assert [r.shape for r in res] == [(b, 2, 2), (b, 2, 2), (b, 3), (b, 2), (b,), (b, 2), (b - 3, 4), (b, 4), (b, 3)]
            assert res[7][0, 0] == 4 * b and res[4].tolist() == list(range(b))
    finally:
        paddle.disable_static()


def test_clone_for_test_switches_random_ops_to_eval():
    """dropout family / rrelu are re-sampled every run of the training program and become identities in `clone(for_test=True)`;
    batch_norm uses its running statistics there."""
    import numpy as np

    import paddle_b200 as paddle

    paddle.enable_static()
    try:
        S = paddle.static
        main, start = S.Program(), S.Program()
        with S.program_guard(main, start):
            x = S.data("x", [4, 6], "float32")
            h = paddle.nn.functional.dropout(x, 0.5, training=True)
            d2 = paddle.nn.Dropout2D(0.5)(paddle.reshape(x, [4, 6, 1, 1]))
            ad = paddle.nn.functional.alpha_dropout(x, 0.5)
            rr = paddle.nn.functional.rrelu(x - 2.0)
            bn = paddle.nn.BatchNorm1D(6)(x * paddle.arange(1, 5).astype("float32").reshape([4, 1]))
        test_prog = main.clone(for_test=True)
        exe = S.Executor()
        exe.run(start)
        X = np.ones((4, 6), "float32")
        a1 = exe.run(main, feed={"x": X}, fetch_list=[h, d2, ad, rr, bn])
        a2 = exe.run(main, feed={"x": X}, fetch_list=[h, d2, ad, rr])
        assert all(not np.allclose(u, v) for u, v in zip(a1[:4], a2))          # fresh masks per run
        assert abs(a1[4].mean()) < 1e-5                                        # training: batch statistics
        b = exe.run(test_prog, feed={"x": X}, fetch_list=[h, d2, ad, rr, bn])
        assert np.allclose(b[0], X) and np.allclose(b[1].reshape(4, 6), X) and np.allclose(b[2], X) and np.allclose(b[3], (X - 2) * (1 / 8 + 1 / 3) / 2)
        assert abs(b[4].mean()) > 0.1                                          # eval: running statistics, not the batch's own
    finally:
        paddle.disable_static()


def test_save_inference_model_prunes_and_pickles_dunder_ops(tmp_path):
    """The saved inference program keeps only what the fetch targets need (the loss sub-graph with `** 2` is dropped), nodes recorded from
    Tensor dunder wrappers survive pickling, and the reloaded program serves any batch size; LR schedulers drive static updates."""
    import os

    import numpy as np

    import paddle_b200 as paddle

    paddle.enable_static()
    try:
        S = paddle.static
        main, start = S.Program(), S.Program()
        with S.program_guard(main, start):
            x = S.data("x", [-1, 3], "float32")
            y = S.data("y", [-1, 1], "float32")
            pred = S.nn.fc(x, 1) ** 2
            loss = paddle.mean((pred - y) ** 2)
            sched = paddle.optimizer.lr.StepDecay(0.05, step_size=2, gamma=0.1)
            paddle.optimizer.SGD(sched).minimize(loss)
        exe = S.Executor()
        exe.run(start)
        X = np.random.RandomState(0).randn(16, 3).astype("float32")
        w = [p for p in main.all_parameters() if p.shape == [3, 1]][0]
        deltas = []
        for i in range(4):
            before = w.numpy().copy()
            exe.run(main, feed={"x": X, "y": X[:, :1]}, fetch_list=[loss])
            deltas.append(float(np.abs(w.numpy() - before).max()))
            sched.step()
        assert deltas[2] < deltas[0] * 0.5                      # lr dropped 10x after two steps
        ref, = exe.run(main.clone(for_test=True), feed={"x": X[:5], "y": X[:5, :1]}, fetch_list=[pred])
        S.save_inference_model(str(tmp_path / "m"), [x], [pred], exe, program=main)
        prog2, feeds, fetches = S.load_inference_model(str(tmp_path / "m"), exe)
        assert len(prog2.nodes) < len(main.nodes) and feeds == ["x"]
        for b in (2, 5):
            out, = exe.run(prog2, feed={"x": X[:b]}, fetch_list=fetches)
            np.testing.assert_allclose(out, ref[:b], rtol=1e-6)
    finally:
        paddle.disable_static()


def test_decomposition_rewrites_composites_into_primitives():
    """paddle.decomposition.decompose: softmax / gelu / layer_norm / silu / mean nodes become primitive nodes; results are unchanged."""
    from paddle_b200 import decomposition as D

    paddle.seed(0)
    main, start = static.Program(), static.Program()
    with static.program_guard(main, start):
        x = static.data("x", [-1, 8], "float32")
        w = paddle.create_parameter([8], "float32", default_initializer=paddle.nn.initializer.Constant(1.5))
        h = paddle.nn.functional.layer_norm(x, [8], weight=w)
        h = paddle.nn.functional.gelu(h) + paddle.nn.functional.silu(x)
        y = paddle.nn.functional.softmax(h, axis=-1)
        z = paddle.nn.functional.log_softmax(h, axis=1).mean()
    exe = static.Executor()
    exe.run(start)
    a = np.random.RandomState(0).rand(5, 8).astype("float32")
    before = exe.run(main, feed={"x": a}, fetch_list=[y, z])
    names = lambda p: [getattr(n.fn, "__name__", str(n.fn)).strip("_") for n in p.nodes]  # noqa: E731
    assert {"softmax", "gelu", "layer_norm", "silu", "log_softmax"} <= set(names(main))
    n_before = len(main.nodes)
    D.decompose(main)
    left = set(names(main)) & {"softmax", "gelu", "layer_norm", "silu", "log_softmax", "mean"}
    assert not left, left
    assert len(main.nodes) > n_before and main.__dict__["_decomposed"] >= 5
    after = exe.run(main, feed={"x": a}, fetch_list=[y, z])
    for p, q in zip(before, after):
        np.testing.assert_allclose(p, q, rtol=2e-5, atol=1e-6)
    # whitelist / blacklist
    m2, s2 = static.Program(), static.Program()
    with static.program_guard(m2, s2):
        x2 = static.data("x", [-1, 8], "float32")
        y2 = paddle.nn.functional.softmax(paddle.nn.functional.gelu(x2), axis=-1)
    D.decompose(m2, blacklist={"gelu"})
    assert "gelu" in names(m2) and "softmax" not in names(m2)
    assert D.has_decomp("rms_norm") and D.get_decomp_rule("softmax") is not None


def test_executor_frees_values_after_their_last_reader():
    """Interpreter GC: the value table holds only what is still needed; results are unchanged; training programs are left alone."""
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import static

    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [4, 8], "float32")
            h = x
            for i in range(12):
                h = paddle.tanh(h * 1.1 + 0.1)
            side = paddle.exp(x)                      # never read, not fetched
            out = h + x
        exe = static.Executor()
        xv = np.random.RandomState(0).randn(4, 8).astype("float32")
        got = exe.run(main, feed={"x": xv}, fetch_list=[out])[0]
        st = exe.last_gc_stats
        assert st is not None and st["freed"] >= 36 and st["peak_live"] <= 6      # 38 values in the program, a handful alive at any time
        ref = xv.copy()
        for i in range(12):
            ref = np.tanh(ref * 1.1 + 0.1)
        np.testing.assert_allclose(got, ref + xv, rtol=1e-5, atol=1e-6)
        mid = exe.run(main, feed={"x": xv}, fetch_list=[h, out])                  # a fetched intermediate survives
        np.testing.assert_allclose(mid[0], ref, rtol=1e-5, atol=1e-6)
        paddle.set_flags({"FLAGS_eager_delete_tensor_gb": -1.0})
        exe.run(main, feed={"x": xv}, fetch_list=[out])
        assert exe.last_gc_stats is None
    finally:
        paddle.set_flags({"FLAGS_eager_delete_tensor_gb": 0.0})
        paddle.disable_static()
    del side


def test_incubate_fused_functionals_are_recorded_in_static_programs():
    """The fused functionals compute on raw tensors; in a program they are recorded as single nodes, so the program follows its feeds instead
    of baking the placeholder values in."""
    import numpy as np
    import torch

    import paddle_b200 as paddle
    import paddle_b200.incubate.nn.functional as IF
    from paddle_b200 import static

    paddle.seed(0)
    w, b = paddle.rand([8]) + 0.5, paddle.randn([8])
    lw, lb = paddle.randn([8, 4]), paddle.randn([4])
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [4, 8], "float32")
            h = IF.fused_rms_norm(x, w, None, 1e-6, 1)
            h = h[0] if isinstance(h, (tuple, list)) else h
            h = IF.fused_layer_norm(h, w, b, 1e-5, begin_norm_axis=1)
            h = h[0] if isinstance(h, (tuple, list)) else h
            h = IF.fused_bias_act(h, b, act_method="gelu")
            y = IF.fused_linear_activation(h, lw, lb, activation="relu") + IF.fused_linear(h, lw, lb) + IF.fused_matmul_bias(h, lw, lb)
        assert len(main.nodes) >= 6
        exe = static.Executor()
        outs = []
        for seed in (1, 2):
            xv = np.random.RandomState(seed).randn(4, 8).astype("float32")
            got = exe.run(main, feed={"x": xv}, fetch_list=[y])[0]
            t = torch.from_numpy(xv)
            wt, bt, lwt, lbt = (v.as_subclass(torch.Tensor) for v in (w, b, lw, lb))
            r = t * torch.rsqrt((t * t).mean(-1, keepdim=True) + 1e-6) * wt
            r = torch.nn.functional.layer_norm(r, (8,), wt, bt, 1e-5)
            r = torch.nn.functional.gelu(r + bt)
            lin = r @ lwt + lbt
            np.testing.assert_allclose(got, (torch.relu(lin) + 2 * lin).numpy(), rtol=2e-4, atol=2e-4)
            outs.append(got)
        assert not np.allclose(outs[0], outs[1])
    finally:
        paddle.disable_static()


def test_raw_tensor_namespaces_are_recorded():
    """geometric / vision.ops / nn.quant functions index raw tensors; in a program each is one recorded node."""
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import static

    idx = paddle.to_tensor([0, 1, 2, 0])
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [4, 8], "float32")
            a = paddle.geometric.send_u_recv(x, idx, idx, "sum")
            b = paddle.geometric.segment_sum(x, paddle.to_tensor([0, 0, 1, 1]))
            r = paddle.vision.ops.roi_align(x.reshape([1, 2, 4, 4]), paddle.to_tensor([[0.0, 0.0, 2.0, 2.0]]), paddle.to_tensor([1], dtype="int32"), 2)
        exe = static.Executor()
        res = []
        for seed in (0, 1):
            xv = np.random.RandomState(seed).randn(4, 8).astype("float32")
            ga, gb, gr = exe.run(main, feed={"x": xv}, fetch_list=[a, b, r])
            ref = np.zeros_like(xv)
            for s, d in zip([0, 1, 2, 0], [0, 1, 2, 0]):
                ref[d] += xv[s]
            np.testing.assert_allclose(ga, ref, rtol=1e-6)
            np.testing.assert_allclose(gb, np.stack([xv[:2].sum(0), xv[2:].sum(0)]), rtol=1e-6)
            res.append(gr)
        assert res[0].shape == (1, 2, 2, 2) and not np.allclose(res[0], res[1])
    finally:
        paddle.disable_static()


def test_host_reads_of_program_variables_raise_while_building():
    """.numpy() / .item() / bool() / int() on a value of the program under construction would hand placeholder zeros to python code; they raise,
    like in the reference.  Constants and fetched results are ordinary tensors."""
    import pytest

    import paddle_b200 as paddle
    from paddle_b200 import static

    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [4, 8], "float32")
            y = (x * 2).sum()
            for read in (lambda: y.numpy(), lambda: y.item(), lambda: bool(y > 0), lambda: float(y), lambda: int(y), lambda: x.tolist()):
                with pytest.raises(RuntimeError, match="no value yet"):
                    read()
            with pytest.raises(RuntimeError, match="no value yet"):
                if y > 0:                                  # value-dependent python branch
                    pass
            assert paddle.to_tensor([3.0]).item() == 3.0   # not a program value
        out = static.Executor().run(main, feed={"x": __import__("numpy").ones((4, 8), "float32")}, fetch_list=[y], return_numpy=False)[0]
        assert float(out) == 64.0 and out.item() == 64.0
    finally:
        paddle.disable_static()


def test_global_scope_reads_and_writes_parameters():
    """`global_scope().find_var(name).get_tensor()`: np.array() reads the live parameter, .set() changes what the next run computes with; a fresh
    Scope is empty; scopes chain through new_scope()."""
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import static

    paddle.enable_static()
    try:
        main, start = static.Program(), static.Program()
        with static.program_guard(main, start):
            x = static.data("x", [2, 4], "float32")
            y = static.nn.fc(x, 3)
        exe = static.Executor()
        exe.run(start)
        wname, bname = [p.name for p in main.all_parameters()]
        sc = static.global_scope()
        w = sc.find_var(wname).get_tensor()
        assert np.array(w).shape == (4, 3) and w.shape() == [4, 3] and w._is_initialized()
        w.set(np.full((4, 3), 0.5, "float32"), paddle.CPUPlace())
        sc.find_var(bname).get_tensor().set(np.array([1.0, 2.0, 3.0], "float32"), paddle.CPUPlace())
        out = exe.run(main, feed={"x": np.ones((2, 4), "float32")}, fetch_list=[y])[0]
        np.testing.assert_allclose(out, [[3.0, 4.0, 5.0]] * 2)
        import pytest

        with pytest.raises(ValueError):
            w.set(np.zeros((2, 2), "float32"), paddle.CPUPlace())
        assert sc.find_var("no_such_var") is None
        with static.scope_guard(static.Scope()):
            assert static.global_scope().find_var(wname) is None
            v = static.global_scope().var("tmp")
            v.get_tensor().set(np.arange(3, dtype="float32"), paddle.CPUPlace())
            assert np.array(static.global_scope().find_var("tmp").get_tensor()).tolist() == [0.0, 1.0, 2.0]
        kid = sc.new_scope()
        assert kid.find_var(wname) is not None and kid.find_var("tmp") is None
    finally:
        paddle.disable_static()


def test_print_py_func_auc_run_with_the_program_and_programs_can_be_inspected(capsys):
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import static

    paddle.enable_static()
    try:
        main, start = static.Program(), static.Program()
        with static.program_guard(main, start):
            x = static.data("x", [2, 4], "float32")
            lbl = static.data("lbl", [2, 1], "int64")
            h = static.nn.fc(x, 2)
            p = static.Print(h, message="h at run", first_n=1)
            q = static.py_func(lambda t: t.numpy().sum() * np.ones((2,), "float32"), [p], None)
            a, ba, _ = static.auc(paddle.nn.functional.softmax(p), lbl)
        assert "h at run" not in capsys.readouterr().out                       # nothing is printed while the program is built
        assert [op.type for op in main.global_block().ops] == ["matmul", "add", "print", "py_func", "softmax", "auc"]
        ops = main.global_block().ops
        assert ops[0].input_arg_names == ["x"] and ops[1].output_arg_names == ops[2].input_arg_names
        text = main.to_string()
        assert "var x : shape[2, 4]" in text and "param fc_w_" in text and "py_func(" in text and str(main) == text
        exe = static.Executor()
        exe.run(start)
        outs = []
        for k in (1.0, 3.0):
            outs.append(exe.run(main, feed={"x": np.ones((2, 4), "float32") * k, "lbl": np.array([[0], [1]])}, fetch_list=[q, h, a]))
        printed = capsys.readouterr().out
        assert printed.count("h at run") == 1 and "shape=[2, 2]" in printed   # first_n = 1
        for q_, h_, a_ in outs:
            np.testing.assert_allclose(q_, np.full((2,), h_.sum()), rtol=1e-6)  # py_func saw the run-time values
            assert 0.0 <= float(a_) <= 1.0
        assert not np.allclose(outs[0][0], outs[1][0])
    finally:
        paddle.disable_static()


def test_static_pylayer_custom_backward_and_row_conv_are_recorded():
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import static

    paddle.enable_static()
    try:
        main, start = static.Program(), static.Program()
        with static.program_guard(main, start):
            x = static.data("x", [3, 4], "float32")
            x.stop_gradient = False
            y = static.nn.static_pylayer(lambda t: paddle.tanh(t), [x], backward_fn=lambda dy: dy * 5.0)       # deliberately "wrong" gradient
            loss = y.sum()
            (gx,) = static.gradients([loss], [x])
            r = static.nn.row_conv(x.reshape([1, 3, 4]), 1)
        exe = static.Executor()
        exe.run(start)
        for seed in (0, 1):
            xv = np.random.RandomState(seed).randn(3, 4).astype("float32")
            yv, gv, rv = exe.run(main, feed={"x": xv}, fetch_list=[y, gx, r])
            np.testing.assert_allclose(yv, np.tanh(xv), rtol=1e-6)
            np.testing.assert_allclose(gv, np.full((3, 4), 5.0), rtol=1e-6)      # the user's backward, not autograd's
            assert rv.shape == (1, 3, 4) and np.abs(rv).sum() > 0
    finally:
        paddle.disable_static()


def test_exponential_moving_average_moves_with_every_run():
    """ema.update() inside a program is an op: EMA_t = d * EMA_{t-1} + (1 - d) * theta_t per Executor.run; apply() swaps in EMA_t / (1 - d^t)."""
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import static

    paddle.enable_static()
    try:
        paddle.seed(0)
        main, start = static.Program(), static.Program()
        with static.program_guard(main, start):
            x = static.data("x", [4, 3], "float32")
            y = static.data("y", [4, 1], "float32")
            loss = ((static.nn.fc(x, 1) - y) ** 2).mean()
            paddle.optimizer.SGD(learning_rate=0.1).minimize(loss)
            ema = static.ExponentialMovingAverage(0.5)
            ema.update()
        exe = static.Executor()
        exe.run(start)
        w = main.all_parameters()[0]
        rng = np.random.RandomState(0)
        feed = {"x": rng.randn(4, 3).astype("float32"), "y": rng.randn(4, 1).astype("float32")}
        ref, hist = np.zeros(tuple(w.shape), np.float32), []
        for t in range(1, 4):
            exe.run(main, feed=feed, fetch_list=[loss])
            cur = np.array(static.global_scope().find_var(w.name).get_tensor())
            hist.append(cur.copy())
            ref = 0.5 * ref + 0.5 * cur                       # the update runs after the optimizer step of the same run
        assert not np.allclose(hist[0], hist[-1])
        with ema.apply(exe):
            inside = np.array(static.global_scope().find_var(w.name).get_tensor())
            np.testing.assert_allclose(inside, ref / (1 - 0.5 ** 3), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(np.array(static.global_scope().find_var(w.name).get_tensor()), hist[-1])      # restored
    finally:
        paddle.disable_static()


def test_dataloader_feed_dicts_dynamic_batch_training_and_the_baked_extent_hint():
    """DataLoader(feed_list=..., return_list=False) yields {name: tensor} feeds; cross_entropy with [-1, 1] labels trains at any batch size (its
    reshape goes by reference to the label); an extent frozen by python code fails with an explanation."""
    import numpy as np
    import pytest

    import paddle_b200 as paddle
    from paddle_b200 import static

    class DS(paddle.io.Dataset):
        def __len__(self):
            return 18

        def __getitem__(self, i):
            return np.random.RandomState(i).randn(4).astype("float32"), np.array([i % 3], "int64")

    paddle.enable_static()
    try:
        paddle.seed(0)
        main, start = static.Program(), static.Program()
        with static.program_guard(main, start):
            x = static.data("x", [-1, 4], "float32")
            y = static.data("y", [-1, 1], "int64")
            loss = paddle.nn.functional.cross_entropy(static.nn.fc(x, 3), y)
            paddle.optimizer.SGD(0.2).minimize(loss)
        loader = paddle.io.DataLoader(DS(), feed_list=[x, y], places=paddle.CPUPlace(), batch_size=4, return_list=False)     # last batch has 2 samples
        exe = static.Executor()
        exe.run(start)
        ls = []
        for _ in range(4):
            for feed in loader:
                assert sorted(feed) == ["x", "y"]
                ls.append(float(exe.run(main, feed=feed, fetch_list=[loss])[0]))
        assert np.mean(ls[-5:]) < np.mean(ls[:5])
        assert np.isfinite(float(exe.run(main, feed=[feed], fetch_list=[loss])[0]))      # a list of per-place dicts is accepted too

        frozen = static.Program()
        with static.program_guard(frozen):
            z = static.data("z", [-1, 4], "float32")
            flat = paddle.reshape(z, [z.size(0) * 4])            # reads the placeholder's extent (1) while building
        assert exe.run(frozen, feed={"z": np.zeros((1, 4), "float32")}, fetch_list=[flat])[0].shape == (4,)
        with pytest.raises(RuntimeError, match="dynamic dims"):
            exe.run(frozen, feed={"z": np.zeros((3, 4), "float32")}, fetch_list=[flat])
    finally:
        paddle.disable_static()


def test_no_public_op_leaks_out_of_a_recorded_program():
    """Sweep: call every public function of paddle.nn.functional / paddle.* that accepts a simple argument pattern on a program variable; its
    tensor results must be values of the program (a body that computes on raw tensors would bake the placeholder's zeros in)."""
    import inspect
    import warnings

    import torch

    import paddle_b200 as paddle
    from paddle_b200 import static

    legit = {"to_tensor", "from_numpy", "get_rng_state", "rank", "get_cuda_rng_state"}          # constants by definition
    skip = {"enable_static", "disable_static", "seed", "set_device", "save", "load", "summary", "flops", "set_flags", "set_default_dtype", "set_grad_enabled",
            "set_printoptions", "manual_seed", "batch", "no_grad", "enable_grad", "set_cuda_rng_state", "set_rng_state", "disable_signal_handler", "check_shape",
            "install_as_paddle"}
    patterns = (((4, 8), lambda x: (x,)), ((1, 2, 4, 4), lambda x: (x,)), ((4, 8), lambda x: (x, x * 0.5 + 0.1)), ((1, 2, 4, 4), lambda x: (x, 2)), ((4, 8), lambda x: (x, 1)))
    leaks, probed = [], 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for label, mod in (("F", paddle.nn.functional), ("paddle", paddle)):
            names = [n for n in getattr(mod, "__all__", dir(mod)) if not n.startswith("_") and hasattr(mod, n)]
            for n in names:
                fn = getattr(mod, n)
                if n in skip or not callable(fn) or inspect.isclass(fn) or inspect.ismodule(fn):
                    continue
                for shape, args in patterns:
                    paddle.enable_static()
                    try:
                        main = static.Program()
                        with static.program_guard(main):
                            x = static.data("x", list(shape), "float32")
                            try:
                                out = fn(*args(x))
                            except BaseException:  # noqa: BLE001  (this argument pattern does not fit the function)
                                continue
                            outs = [o for o in (out if isinstance(out, (list, tuple)) else [out]) if isinstance(o, torch.Tensor)]
                            if not outs:
                                continue
                            probed += 1
                            if n not in legit and not all(id(o) in main._fetch_alias for o in outs):
                                leaks.append(f"{label}.{n}")
                            break
                    finally:
                        paddle.disable_static()
    assert probed > 350 and leaks == [], leaks


def test_samplers_draw_anew_at_every_run_of_a_program():
    """rand / randn / randint / uniform / normal / randperm / bernoulli / multinomial inside a program are ops, not constants: two runs differ, a
    re-seeded run repeats."""
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import static

    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [4, 8], "float32")
            outs = [x + paddle.rand([4, 8]), x + paddle.randn([4, 8]), x + paddle.randint(0, 1000, [4, 8]).astype("float32"), x + paddle.uniform([4, 8]),
                    x + paddle.normal(0.0, 1.0, [4, 8]), x[0] + paddle.randperm(8).astype("float32"), x + paddle.bernoulli(paddle.full([4, 8], 0.5)),
                    x[0, :5] + paddle.multinomial(paddle.ones([50]), 5).astype("float32")]
        exe = static.Executor()
        xv = np.zeros((4, 8), "float32")
        a = exe.run(main, feed={"x": xv}, fetch_list=outs)
        b = exe.run(main, feed={"x": xv}, fetch_list=outs)
        assert all(not np.allclose(p, q) for p, q in zip(a, b))
        paddle.seed(11)
        c = exe.run(main, feed={"x": xv}, fetch_list=outs)
        paddle.seed(11)
        d = exe.run(main, feed={"x": xv}, fetch_list=outs)
        assert all(np.allclose(p, q) for p, q in zip(c, d))
    finally:
        paddle.disable_static()


def test_distributions_in_a_program_follow_the_feeds_and_resample():
    """Distribution objects wrap torch.distributions built from raw tensors; in a program a method call is one recorded node that rebuilds the
    distribution from its constructor arguments at run time (nested distributions included)."""
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import static

    D = paddle.distribution
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [4], "float32")
            free = D.Normal(0.0, 1.0).sample([4]) + x                       # nothing of the program feeds it: still one draw per run
            d = D.Normal(x, 1.0)
            outs = [free, d.sample([2]), D.Categorical(paddle.ones([50])).sample([4]).astype("float32") + x, d.log_prob(paddle.zeros([4])), d.mean,
                    D.kl_divergence(d, D.Normal(0.0, 1.0)), D.Independent(D.Normal(x, 1.0), 1).log_prob(paddle.zeros([4])), d.entropy()]
        exe = static.Executor()
        xv = np.arange(4, dtype="float32")
        a = exe.run(main, feed={"x": xv}, fetch_list=outs)
        b = exe.run(main, feed={"x": xv}, fetch_list=outs)
        assert all(not np.allclose(p, q) for p, q in zip(a[:3], b[:3]))       # samples
        lp = -0.5 * xv ** 2 - 0.5 * np.log(2 * np.pi)
        np.testing.assert_allclose(a[3], lp, atol=1e-5)
        np.testing.assert_allclose(a[4], xv)
        np.testing.assert_allclose(a[5], 0.5 * xv ** 2, atol=1e-5)
        np.testing.assert_allclose(a[6], lp.sum(), atol=1e-4)
        np.testing.assert_allclose(a[7], np.full(4, 0.5 * np.log(2 * np.pi * np.e)), atol=1e-5)
    finally:
        paddle.disable_static()
    n = D.Normal(0.0, 2.0)                                                    # dynamic mode: unchanged
    assert n.sample([3]).shape == [3] and abs(float(n.entropy()) - 0.5 * np.log(2 * np.pi * np.e * 4)) < 1e-5


def test_programs_with_samplers_and_distribution_nodes_save_and_load(tmp_path):
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import static

    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [-1, 4], "float32")
            y = x + paddle.randn([4]) + paddle.distribution.Normal(0.0, 1.0).sample([4])
            y = paddle.nn.functional.interpolate(y.reshape([1, 1, -1, 4]), scale_factor=2).reshape([-1, 8])
        exe = static.Executor()
        static.save_inference_model(str(tmp_path / "m"), [x], [y], exe, program=main)
        prog, feeds, fetch = static.load_inference_model(str(tmp_path / "m"), exe)
        feed = {feeds[0]: np.ones((3, 4), "float32")}
        a, b = exe.run(prog, feed=feed, fetch_list=fetch)[0], exe.run(prog, feed=feed, fetch_list=fetch)[0]
        assert a.shape == (6, 8) and not np.allclose(a, b)          # dynamic batch, per-run sampling survive the round trip
    finally:
        paddle.disable_static()
