"""API-surface fillers: fleet utils / Role / UtilBase / data generators, sparse 2-D conv, distribution.transform, misc."""
import numpy as np
import pytest

import paddle_b200 as paddle


def test_fleet_utils_and_roles(tmp_path):
    from paddle_b200.distributed import fleet
    from paddle_b200.distributed.fleet.utils import hybrid_parallel_util, sequence_parallel_utils, tensor_fusion_helper

    assert fleet.Role.SERVER == 2
    u = fleet.UtilBase()
    assert np.allclose(u.all_reduce([1, 2]), [1, 2]) and u.all_gather(5) == [5]
    assert u.get_file_shard(["a", "b", "c"]) == ["a", "b", "c"]
    fs = fleet.utils.LocalFS()
    d = str(tmp_path / "x")
    fs.mkdirs(d)
    fs.touch(d + "/f")
    assert fs.ls_dir(str(tmp_path)) == (["x"], []) and fs.is_file(d + "/f")
    fs.mv(d + "/f", d + "/g")
    assert fs.is_exist(d + "/g")
    fs.delete(d)
    assert not fs.is_exist(d)
    with pytest.raises(RuntimeError):
        fleet.utils.HDFSClient("/none", {}).mkdirs("/x")
    assert fleet.utils.recompute is fleet.recompute
    assert hasattr(sequence_parallel_utils, "ColumnSequenceParallelLinear") and hasattr(hybrid_parallel_util, "fused_allreduce_gradients")
    lin = paddle.nn.Linear(4, 4)
    decay, fused, buckets = tensor_fusion_helper.fused_parameters(lin.parameters())
    assert sum(int(t.numel()) for t in fused) >= 20 and buckets


def test_data_generator():
    from paddle_b200.distributed import fleet

    class G(fleet.MultiSlotDataGenerator):
        def generate_sample(self, line):
            def it():
                yield [("words", [1, 2, 3]), ("label", [0])]
            return it

    assert G().run_from_memory() == ["3 1 2 3 1 0\n"]


def test_sparse_conv2d():
    idx = np.array([[0, 0], [1, 2], [1, 3]])
    sp = paddle.sparse.sparse_coo_tensor(idx, np.random.rand(2, 3).astype("float32"), [1, 4, 4, 3])
    sub = paddle.sparse.nn.SubmConv2D(3, 5, 3)
    y = sub(sp)
    assert list(y.shape) == [1, 4, 4, 5] and y.nnz() == 2
    conv = paddle.sparse.nn.Conv2D(3, 5, 3)
    z = conv(sp)
    assert list(z.shape) == [1, 2, 2, 5]
    import torch
    import torch.nn.functional as F

    ref = F.conv2d(sp.to_dense().permute(0, 3, 1, 2), conv.weight.permute(3, 2, 0, 1), conv.bias).permute(0, 2, 3, 1)
    zd = z.to_dense()
    mask = zd.abs().sum(-1) > 0
    assert torch.allclose(torch.as_tensor(zd)[mask], torch.as_tensor(ref)[mask], atol=1e-5)
    assert paddle.sparse.nn.functional.subm_conv2d_igemm(sp, sub.weight, sub.bias).nnz() == 2


def test_misc_fillers():
    import paddle_b200.distribution.transform as T

    assert T.ExpTransform is paddle.distribution.ExpTransform
    assert paddle.nn.quant.Stub()(3) == 3
    assert paddle.inference._get_phi_kernel_name("elementwise_add") == "add"
    assert paddle.incubate.optimizer.LBFGS is paddle.optimizer.LBFGS

    class L(paddle.nn.Layer):
        @paddle.amp.debugging.check_layer_numerics
        def forward(self, x):
            return x / 0.0

    with pytest.raises(RuntimeError):
        L()(paddle.ones([2]))


def test_distributed_passes_and_communication():
    import paddle_b200.distributed.communication as comm
    from paddle_b200.distributed.passes import PassContext, PassManager, new_pass

    assert comm.all_reduce is paddle.distributed.all_reduce and hasattr(comm.group, "new_group") and hasattr(comm.stream, "all_reduce")
    paddle.enable_static()
    try:
        main = paddle.static.Program()
        with paddle.static.program_guard(main):
            x = paddle.static.data("x", [4, 8], "float32")
            w = paddle.static.create_parameter([8, 8], "float32")
            y = paddle.matmul(x, w)
            z = paddle.nn.functional.relu(y)
        pm = PassManager([new_pass("auto_parallel_bf16"), new_pass("auto_parallel_gradient_merge_pass", {"k_steps": 4})], context=PassContext())
        ctx = pm.apply([main])
        assert main._dist_attrs["amp"]["wrapped"] >= 1 and main._dist_attrs["gradient_merge"]["k_steps"] == 4 and len(ctx.passes) == 2
        exe = paddle.static.Executor()
        out, = exe.run(main, feed={"x": np.ones((4, 8), "float32")}, fetch_list=[z])
        assert out.shape == (4, 8)
    finally:
        paddle.disable_static()


def test_elastic_manager_membership():
    import time

    from paddle_b200.distributed.fleet import elastic

    a = elastic.ElasticManager(np="1:2", host="a", heartbeat_s=0.05, ttl_s=0.3).start_heartbeat()
    assert a.wait(2) and a.status() == "running"
    b = elastic.ElasticManager(np="1:2", host="b", store=a.store, heartbeat_s=0.05, ttl_s=0.3).start_heartbeat()
    time.sleep(0.15)
    assert a.status() == elastic.ElasticStatus.RESTART and a.hosts() == ["a", "b"]
    b.exit()
    time.sleep(0.5)
    assert a.hosts() == ["a"]
    a.exit()
    c = elastic.ElasticManager(np="2", host="c", heartbeat_s=0.05, ttl_s=0.3).start_heartbeat()
    assert c.status() == elastic.ElasticStatus.HOLD and not c.wait(0.3)
    c.exit()


def test_tensorrt_convert_api(tmp_path):
    import paddle_b200.tensorrt as trt

    net = paddle.nn.Sequential(paddle.nn.Linear(8, 16), paddle.nn.ReLU(), paddle.nn.Linear(16, 4))
    cfg = trt.TensorRTConfig([trt.Input(min_input_shape=(1, 8), optim_input_shape=(4, 8), max_input_shape=(8, 8))], precision_mode=trt.PrecisionMode.FP32)
    fn = trt.convert_loaded_model(net, cfg)
    x = paddle.ones([4, 8])
    assert np.allclose(fn(x).numpy(), net(x).numpy(), atol=1e-6)
    prefix = str(tmp_path / "m")
    paddle.jit.save(net, prefix, input_spec=[paddle.static.InputSpec([None, 8], "float32")])
    cfg.save_model_dir = str(tmp_path / "out" / "m")
    fn2 = trt.convert(prefix, cfg)
    assert np.allclose(fn2(x).numpy(), net(x).numpy(), atol=1e-5)
    assert np.allclose(paddle.jit.load(cfg.save_model_dir)(x).numpy(), net(x).numpy(), atol=1e-5)


def test_native_tracer_and_profiler(tmp_path):
    """C++ range tracer (csrc/runtime/tracer.cpp): nesting, threads, chrome export; wired into profiler.Profiler."""
    import json
    import threading

    from paddle_b200 import _build, profiler

    C = _build.load(required=False)
    if C is None:
        pytest.skip("native extension not built")
    C.tracer_collect()
    C.tracer_enable(1)
    C.tracer_begin("outer", 0)
    C.tracer_begin("inner", 1)
    C.tracer_end()
    C.tracer_end()
    t = threading.Thread(target=lambda: (C.tracer_begin("thr", 2), C.tracer_end()))
    t.start()
    t.join()
    evs = {e[0]: e for e in C.tracer_collect()}
    assert set(evs) == {"outer", "inner", "thr"} and evs["inner"][3] == 1 and evs["outer"][3] == 0
    assert evs["outer"][4] <= evs["inner"][4] <= evs["inner"][5] <= evs["outer"][5] and evs["thr"][2] != evs["outer"][2]
    C.tracer_begin('we"ird', 0)
    C.tracer_end()
    p = str(tmp_path / "t.json")
    assert C.tracer_export_chrome(p, 1) == 1 and json.load(open(p))["traceEvents"][0]["name"] == 'we"ird'
    C.tracer_enable(0)
    C.tracer_begin("ignored", 0)
    C.tracer_end()
    assert C.tracer_collect() == []

    with profiler.Profiler(targets=[profiler.ProfilerTarget.CPU]) as prof:
        with profiler.RecordEvent("user"):
            C.tracer_begin("fake_kernel", 1)
            C.tracer_end()
        prof.step()
    stats = profiler.kernel_statistics(prof)
    assert stats["fake_kernel"][0] == 1
    out = str(tmp_path / "trace.json")
    prof.export(out)
    names = {e.get("name") for e in json.load(open(out))["traceEvents"]}
    assert "fake_kernel" in names and "user" in names


def test_native_best_fit_allocator():
    """csrc/runtime/offset_allocator.h (backs the symmetric heap): alignment, reuse, coalescing, no overlap under churn."""
    import random

    from paddle_b200 import _build

    C = _build.load(required=False)
    if C is None:
        pytest.skip("native extension not built")
    cap = 1 << 20
    a = C.BestFitAllocator(1024, cap)
    o1, o2, o3 = a.alloc(1000, 1024), a.alloc(5000, 1024), a.alloc(300)
    assert (o1, o2) == (1024, 2048) and a.block_size(o2) == 5120 and a.in_use() == 1024 + 5120 + 512
    a.free(o2)
    assert a.num_free_blocks() == 2 and a.alloc(4096, 1024) == o2   # best fit reuses the hole
    rng, live = random.Random(0), {}
    for _ in range(3000):
        if live and rng.random() < 0.45:
            k = rng.choice(list(live))
            a.free(k)
            del live[k]
            continue
        n, al = rng.randint(1, 20000), rng.choice([256, 1024, 4096])
        try:
            o = a.alloc(n, al)
        except RuntimeError:
            continue
        assert o % al == 0 and o >= 1024 and o + n <= cap
        assert all(o + n <= k or k + v <= o for k, v in live.items())
        live[o] = n
    for k in live:
        a.free(k)
    a.release_from(0)
    assert a.in_use() == 0 and a.num_free_blocks() == 1 and a.largest_free() == cap - 1024 and a.peak() > 0
    with pytest.raises(RuntimeError, match="out of memory"):
        a.alloc(cap * 2)
    with pytest.raises(RuntimeError):
        a.free(12345)


def test_multislot_dataset_native_feed(tmp_path):
    """InMemoryDataset / QueueDataset over the native multi-slot feed (csrc/runtime/data_feed.cpp), fed by MultiSlotDataGenerator."""
    from paddle_b200 import _build
    from paddle_b200.distributed import InMemoryDataset, QueueDataset, fleet

    if _build.load(required=False) is None:
        pytest.skip("native extension not built")

    class Gen(fleet.MultiSlotDataGenerator):
        def __init__(self, base):
            super().__init__()
            self.base = base

        def generate_sample(self, line):
            def it():
                for i in range(5):
                    k = self.base + i
                    yield [("ids", list(range(k, k + 1 + i % 3))), ("dense", [k * 0.5, k * 0.25]), ("label", [k % 2])]
            return it

    files = []
    for j, base in enumerate((0, 100)):
        p = tmp_path / f"part-{j}"
        p.write_text("".join(Gen(base).run_from_memory()))
        files.append(str(p))
    ds = InMemoryDataset()
    ds.init(batch_size=4, thread_num=2, use_var=[("ids", "int64"), ("dense", "float32"), ("label", "int64")])
    ds.set_filelist(files)
    ds.load_into_memory()
    assert ds.get_memory_data_size() == 10
    batches = list(ds)
    assert [len(b["label"]) for b in batches] == [4, 4, 2]
    b0 = batches[0]
    vals, lod = b0["ids"]                       # ragged slot -> (values, lod)
    assert list(lod) == [0, 1, 3, 6, 7] and list(vals[:3]) == [0, 1, 2]
    assert b0["dense"].shape == (4, 2) and np.allclose(b0["dense"][1], [0.5, 0.25]) and b0["label"].shape == (4, 1)
    ds.local_shuffle(seed=3)
    labels = np.concatenate([b["dense"][:, 0] for b in ds])
    assert sorted(labels) == sorted(k * 0.5 for k in list(range(5)) + list(range(100, 105))) and list(labels) != sorted(labels)
    ds.global_shuffle()
    assert ds.get_shuffle_data_size() == 10
    ds.release_memory()
    assert ds.get_memory_data_size() == 0

    q = QueueDataset()
    q.init(batch_size=3, use_var=[("ids", "int64"), ("dense", "float32"), ("label", "int64")])
    q.set_filelist(files)
    assert sum(len(b["label"]) for b in q) == 10
    bad = tmp_path / "bad"
    bad.write_text("2 1\n")
    ds.set_filelist([str(bad)])
    with pytest.raises(RuntimeError, match="line 1"):
        ds.load_into_memory()


def test_capture_train_step_cpu_falls_back_to_eager():
    paddle.seed(3)
    net = paddle.nn.Linear(4, 2)
    opt = paddle.optimizer.AdamW(1e-2, parameters=net.parameters())
    opt.enable_flat_arena()
    step = paddle.jit.capture_train_step(lambda x: (net(x) ** 2).mean(), opt, warmup=1)
    x = paddle.ones([3, 4])
    losses = [float(step(x)) for _ in range(5)]
    assert not step.captured and losses[-1] < losses[0] and opt._step_count == 5


def test_quasi_newton_minimizers_asp_module_groupwise_observer():
    from paddle_b200.incubate.optimizer.functional import minimize_bfgs, minimize_lbfgs

    def rosen(x):
        return (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2

    conv, calls, pos, val, grad, H = minimize_bfgs(rosen, paddle.to_tensor([-1.2, 1.0]), max_iters=200, dtype="float64")
    assert bool(conv) and np.allclose(pos.numpy(), [1, 1], atol=1e-5) and float(val) < 1e-10 and H.shape == [2, 2] and int(calls) > 10
    conv, calls, pos, val, grad = minimize_lbfgs(rosen, paddle.to_tensor([-1.2, 1.0]), max_iters=200, dtype="float64", history_size=5)
    assert bool(conv) and np.allclose(pos.numpy(), [1, 1], atol=1e-5) and float(grad.abs().max()) < 1e-5
    import paddle_b200.incubate.asp as asp

    class MyProj(paddle.nn.Layer):
        def __init__(self):
            super().__init__()
            self.weight = self.create_parameter([8, 8])

        def forward(self, x):
            return x @ self.weight

    asp.add_supported_layer(MyProj)
    m = MyProj()
    asp.prune_model(m)
    assert abs(asp.calculate_density(m.weight) - 0.5) < 1e-6
    from paddle_b200.quantization.observers import GroupWiseWeightObserver, GroupWiseWeightObserverLayer

    ob = GroupWiseWeightObserverLayer(None, quant_bits=4, group_size=4)
    w = paddle.to_tensor(np.arange(32, dtype="float32").reshape(8, 4) - 10)
    ob(w)
    assert ob.scales().shape == [2, 4] and float(ob.scales()[1, 3]) == 21.0 and ob.bit_length() == 4 and GroupWiseWeightObserver(4, 4) is not None


def test_reference_signatures_spot_checks():
    """Argument names / order that differ from a naive guess (found by an AST scan of the reference sources)."""
    import inspect

    import paddle_b200.distributed as dist

    def names(f):
        return list(inspect.signature(f).parameters)

    assert names(dist.alltoall)[:2] == ["in_tensor_list", "out_tensor_list"] and names(dist.alltoall_single)[:2] == ["in_tensor", "out_tensor"]
    assert names(dist.stream.alltoall)[:2] == ["out_tensor_or_tensor_list", "in_tensor_or_tensor_list"] and "use_calc_stream" in names(dist.stream.all_reduce)
    assert names(dist.get_group) == ["id"] and names(dist.unshard_dtensor) == ["dist_tensor"]
    assert names(paddle.nn.Conv2DTranspose.__init__)[7:9] == ["dilation", "groups"] and names(paddle.nn.Conv1DTranspose.__init__)[7:9] == ["groups", "dilation"]
    assert names(paddle.nn.SimpleRNN.__init__)[7] == "activation" and "proj_size" in names(paddle.nn.LSTM.__init__) and "proj_size" not in names(paddle.nn.GRU.__init__)
    assert names(paddle.nn.Embedding.__init__)[7] == "weight_attr" and names(paddle.nn.SpectralNorm.__init__)[4] == "eps"
    assert names(paddle.round)[:2] == ["x", "decimals"] and names(paddle.t)[0] == "input" and names(paddle.seed) == ["seed"]
    assert names(paddle.nn.functional.temporal_shift)[3:] == ["name", "data_format"]
    I = paddle.incubate.nn.functional
    assert names(I.fused_moe)[:4] == ["x", "gate_weight", "ffn1_weight", "ffn2_weight"] and names(I.fused_dot_product_attention)[:5] == ["query", "key", "value", "attn_mask", "dropout_p"]
    assert paddle.utils.flops("matmul_v2", {"X": [[4, 8]], "Y": [[8, 16]]}, {}) == 2 * 4 * 8 * 16
    assert paddle.utils.flops("conv2d", {"Input": [[1, 3, 8, 8]], "Filter": [[4, 3, 3, 3]]}, {"strides": [1, 1], "paddings": [1, 1], "dilations": [1, 1]}) == 2 * 4 * 8 * 8 * 27
    x = paddle.to_tensor(np.array([-8, 8, -1], "int32"))
    y = paddle.to_tensor(np.array([1, 2, 31], "int32"))
    assert paddle.bitwise_right_shift(x, y, is_arithmetic=False).numpy().tolist() == [2147483644, 2, 1]
    assert np.allclose(paddle.round(paddle.to_tensor([1.2345, -2.555]), decimals=2).numpy(), [1.23, -2.56], atol=1e-6)


def test_class_method_parity_additions(tmp_path):
    net = paddle.nn.Sequential(paddle.nn.Linear(3, 3))
    shared = paddle.nn.Linear(2, 2)

    class Twice(paddle.nn.Layer):
        def __init__(self):
            super().__init__()
            self.a, self.b = shared, shared

    t = Twice()
    assert len(list(t.named_parameters())) == 2 and len(list(t.named_parameters(remove_duplicate=False))) == 4
    assert len(list(t.named_sublayers(remove_duplicate=False))) == 2
    h = net.register_state_dict_hook(lambda sd: {k.upper(): v for k, v in sd.items()})
    assert all(k.isupper() or not k.isalpha() for k in net.state_dict()) and "0.WEIGHT" in net.state_dict()
    h.remove()
    assert "0.weight" in net.state_dict()
    net.append(module=paddle.nn.ReLU()).insert(0, module=paddle.nn.Tanh()).extend(sequential=[paddle.nn.Sigmoid()])
    assert len(net) == 4
    opt = paddle.optimizer.SGD(0.1, parameters=net.parameters())
    loss = net(paddle.ones([2, 3])).sum()
    pg = opt.backward(loss)
    assert len(pg) == 2 and all(g is not None for _, g in pg)
    before = net[1].weight.numpy().copy()
    reg = opt.append_regularization_ops(pg, paddle.regularizer.L2Decay(0.5))
    assert np.allclose(reg[0][1].numpy(), pg[0][1].numpy() + 0.5 * pg[0][0].numpy(), atol=1e-6)
    opt.apply_gradients(pg)
    assert not np.allclose(net[1].weight.numpy(), before)
    sc = paddle.amp.GradScaler()
    sc.set_incr_ratio(3.0)
    sc.set_decr_ratio(0.25)
    sc.set_incr_every_n_steps(7)
    sc.set_decr_every_n_nan_or_inf(2)
    sc.set_init_loss_scaling(new_init_loss_scaling=128.0)
    assert (sc.get_incr_ratio(), sc.get_decr_ratio(), sc.get_incr_every_n_steps(), sc.get_decr_every_n_nan_or_inf(), sc.get_init_loss_scaling()) == (3.0, 0.25, 7, 2, 128.0)
    st = paddle.distributed.fleet.DistributedStrategy()
    st.qat_configs = {"weight_bits": 4}
    st.save_to_prototxt(str(tmp_path / "s.txt"))
    st2 = paddle.distributed.fleet.DistributedStrategy()
    st2.load_from_prototxt(str(tmp_path / "s.txt"))
    assert st2.qat_configs["weight_bits"] == 4 and st2.sync_batch_norm is False and st2.localsgd_configs["k_steps"] == 1


def test_is_sparse_method_and_attribute_spellings():
    import copy

    x = paddle.to_tensor(np.eye(2, dtype="float32"))
    assert x.is_sparse() is False and not x.is_sparse and x.is_dense() and repr(x.is_sparse) == "False"
    assert copy.deepcopy(x).shape == [2, 2]
    sp = x.to_sparse_coo(2)
    assert sp.is_sparse_coo() and bool(sp.is_sparse) and sp.nnz() == 2      # sparse results are torch sparse tensors with paddle methods patched on
