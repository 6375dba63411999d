"""Worker body for the multi-process tests. Each case compares the parallel result with a single-process reference
computed locally with identical seeds (the reference's hybrid_parallel_* test pattern)."""
import os
import sys

import numpy as np
import torch

import paddle_b200 as paddle
import paddle_b200.distributed as dist
from paddle_b200 import nn
from paddle_b200.distributed import fleet

GPU = torch.cuda.is_available() and os.environ.get("B200_TEST_GPU", "0") == "1"


def setup(dp=1, mp=1, pp=1, sharding=1):
    if GPU:
        paddle.set_device(f"gpu:{os.environ['LOCAL_RANK']}")
    s = fleet.DistributedStrategy()
    s.hybrid_configs = {"dp_degree": dp, "mp_degree": mp, "pp_degree": pp, "sharding_degree": sharding}
    fleet.init(is_collective=True, strategy=s)
    return s, fleet.get_hybrid_communicate_group()


def close(a, b, tol=1e-4):
    a, b = torch.as_tensor(np.asarray(a, dtype=np.float64)), torch.as_tensor(np.asarray(b, dtype=np.float64))
    err = (a - b).abs().max().item() / max(1e-8, b.abs().max().item())
    assert err < tol, f"mismatch {err}"


def case_collectives():
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    t = paddle.to_tensor([float(r + 1)] * 4)
    dist.all_reduce(t)
    close(t.numpy(), [sum(range(1, w + 1))] * 4)
    outs = []
    dist.all_gather(outs, paddle.to_tensor([float(r)]))
    close([o.item() for o in outs], list(range(w)))
    b = paddle.to_tensor([float(r)])
    dist.broadcast(b, 1)
    close(b.numpy(), [1.0])
    objs = []
    dist.all_gather_object(objs, {"r": r})
    assert [o["r"] for o in objs] == list(range(w))
    ins = [paddle.to_tensor([float(r * 10 + j)]) for j in range(w)]
    res = []
    dist.alltoall(ins, res)
    close([x.item() for x in res], [j * 10 + r for j in range(w)])
    res2 = []
    dist.stream.alltoall(res2, ins)          # the stream variant takes the output first
    close([x.item() for x in res2], [j * 10 + r for j in range(w)])
    g = dist.new_group(list(range(w)))
    x = paddle.to_tensor([1.0])
    dist.all_reduce(x, group=g)
    close(x.numpy(), [float(w)])
    if r == 0:
        dist.send(paddle.to_tensor([42.0]), dst=1)
    elif r == 1:
        y = paddle.zeros([1])
        dist.recv(y, src=0)
        close(y.numpy(), [42.0])
    dist.barrier()


def _mlp_ref(seed, din, dh, dout):
    paddle.seed(seed)
    return nn.Linear(din, dh), nn.Linear(dh, dout)


def case_mp_layers():
    """Column+Row parallel MLP == dense MLP (loss and grads). Parity: test/collective/fleet/hybrid_parallel_mp_layers.py."""
    _, hcg = setup(mp=2)
    r = hcg.get_model_parallel_rank()
    paddle.seed(7)
    w1 = paddle.randn([16, 32]) * 0.1
    w2 = paddle.randn([32, 8]) * 0.1
    x = paddle.randn([4, 16])
    col = fleet.ColumnParallelLinear(16, 32, has_bias=False, gather_output=False)
    row = fleet.RowParallelLinear(32, 8, has_bias=False, input_is_parallel=True)
    col.weight.set_value(w1[:, r * 16:(r + 1) * 16])
    row.weight.set_value(w2[r * 16:(r + 1) * 16])
    y = row(paddle.nn.functional.relu(col(x)))
    loss = (y * y).sum()
    loss.backward()
    w1r, w2r = paddle.to_tensor(w1, stop_gradient=False), paddle.to_tensor(w2, stop_gradient=False)
    yr = paddle.nn.functional.relu(x @ w1r) @ w2r
    lr = (yr * yr).sum()
    lr.backward()
    close(loss.item(), lr.item())
    close(col.weight.grad.numpy(), w1r.grad[:, r * 16:(r + 1) * 16].numpy())
    close(row.weight.grad.numpy(), w2r.grad[r * 16:(r + 1) * 16].numpy())
    # vocab parallel embedding + parallel cross entropy
    emb = fleet.VocabParallelEmbedding(20, 6)
    full = paddle.randn([20, 6])
    emb.weight.set_value(full[r * 10:(r + 1) * 10])
    ids = paddle.to_tensor([[1, 15, 7], [19, 0, 11]])
    close(emb(ids).numpy(), full[ids].numpy())
    logits = paddle.randn([5, 12])
    lab = paddle.to_tensor([0, 3, 11, 6, 7])
    pce = fleet.ParallelCrossEntropy()
    lp = pce(logits[:, r * 6:(r + 1) * 6], lab)
    ref = paddle.nn.functional.cross_entropy(logits, lab, reduction="none")
    close(lp.numpy().reshape(-1), ref.numpy().reshape(-1))


def case_sequence_parallel():
    _, hcg = setup(mp=2)
    r = hcg.get_model_parallel_rank()
    paddle.seed(3)
    w1 = paddle.randn([8, 16]) * 0.2
    w2 = paddle.randn([16, 8]) * 0.2
    x = paddle.randn([6, 2, 8])  # [S, B, H]
    col = fleet.ColumnSequenceParallelLinear(8, 16, has_bias=False)
    row = fleet.RowSequenceParallelLinear(16, 8, has_bias=False)
    col.weight.set_value(w1[:, r * 8:(r + 1) * 8])
    row.weight.set_value(w2[r * 8:(r + 1) * 8])
    xl = paddle.to_tensor(x[r * 3:(r + 1) * 3], stop_gradient=False)
    y = row(paddle.tanh(col(xl)))
    (y * y).sum().backward()
    xr = paddle.to_tensor(x, stop_gradient=False)
    w1r, w2r = paddle.to_tensor(w1, stop_gradient=False), paddle.to_tensor(w2, stop_gradient=False)
    yr = paddle.tanh(xr @ w1r) @ w2r
    (yr * yr).sum().backward()
    close(y.numpy(), yr[r * 3:(r + 1) * 3].numpy())
    close(xl.grad.numpy(), xr.grad[r * 3:(r + 1) * 3].numpy())
    close(col.weight.grad.numpy(), w1r.grad[:, r * 8:(r + 1) * 8].numpy())
    close(row.weight.grad.numpy(), w2r.grad[r * 8:(r + 1) * 8].numpy())


def _tiny_llama(cfg_kw):
    from paddle_b200.models import llama as L

    return L, L.llama_tiny(dtype="float32", **cfg_kw)


def _train_ref(L, cfg, ids, steps, lr=1e-2):
    paddle.seed(11)
    m = L.LlamaForCausalLM(cfg)
    opt = paddle.optimizer.AdamW(lr, parameters=m.parameters(), weight_decay=0.0)
    losses = []
    for _ in range(steps):
        loss = m(ids[:, :-1], ids[:, 1:])
        loss.backward()
        opt.step()
        opt.clear_grad()
        losses.append(loss.item())
    return m, losses


def case_dp():
    """DataParallel: 2 ranks x half batch == 1 rank x full batch. Parity: test/collective parallel_dygraph_*."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    paddle.seed(5)
    net = nn.Sequential(nn.Linear(8, 16), nn.Tanh(), nn.Linear(16, 4))
    ref = nn.Sequential(nn.Linear(8, 16), nn.Tanh(), nn.Linear(16, 4))
    ref.set_state_dict(net.state_dict())
    x = paddle.randn([4 * w, 8])
    y = paddle.randn([4 * w, 4])
    dp = paddle.DataParallel(net, comm_buffer_size=1)
    opt = paddle.optimizer.SGD(0.1, parameters=dp.parameters())
    ropt = paddle.optimizer.SGD(0.1, parameters=ref.parameters())
    for _ in range(3):
        sl = slice(r * 4, (r + 1) * 4)
        loss = ((dp(x[sl]) - y[sl]) ** 2).mean()
        loss.backward()
        opt.step()
        opt.clear_grad()
        lr_ = ((ref(x) - y) ** 2).mean()
        lr_.backward()
        ropt.step()
        ropt.clear_grad()
    for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
        close(a.numpy(), b.numpy(), 1e-4)


def case_pp():
    """PipelineParallel 1F1B (pp=2, 4 micro-batches) == single-process training. Parity: hybrid_parallel_pp_*.py."""
    s, hcg = setup(pp=2)
    L, cfg = _tiny_llama({})
    ids = paddle.to_tensor(np.random.RandomState(0).randint(0, cfg.vocab_size, (4, 33)))
    s.pipeline_configs = {"accumulate_steps": 4, "micro_batch_size": 1, "schedule_mode": os.environ.get("B200_TEST_PP_MODE", "1F1B")}
    paddle.seed(11)
    from paddle_b200.distributed.fleet.pipeline import PipelineLayer

    # build the reference first with the same seed stream, then load its weights into this stage's layers
    ref, ref_losses = _train_ref(L, cfg, ids, 0)
    ref_sd = {k: v.clone() for k, v in ref.state_dict().items()}
    pl = PipelineLayer(L.pipeline_layer_descs(cfg), num_stages=2, loss_fn=L.LlamaPretrainingCriterion(cfg), seg_method="layer:LlamaDecoderLayer")
    # map pipeline names -> reference names
    mapping = {}
    for name, _ in pl.named_parameters():
        idx, rest = name.split(".", 1)
        i = int(idx)
        if i == 0:
            mapping[name] = "llama.embedding." + rest
        elif i == cfg.num_hidden_layers + 1:
            mapping[name] = "lm_head." + rest
        else:
            mapping[name] = f"llama.layers.{i - 1}." + rest
    pl.set_state_dict({k: ref_sd[v] for k, v in mapping.items()})
    model = fleet.distributed_model(pl)
    opt = fleet.distributed_optimizer(paddle.optimizer.AdamW(1e-2, parameters=pl.parameters(), weight_decay=0.0))
    ropt = paddle.optimizer.AdamW(1e-2, parameters=ref.parameters(), weight_decay=0.0)
    if os.environ.get("B200_TEST_PP_MODE", "1F1B").upper().startswith("ZB"):
        sched = model.get_static_scheduler()
        assert "w0" in sched and type(model).__name__ == "PipelineParallelZeroBubble", sched
    for _ in range(3):
        loss = model.train_batch([ids[:, :-1], ids[:, 1:]], opt)
        rl = ref(ids[:, :-1], ids[:, 1:])
        rl.backward()
        ropt.step()
        ropt.clear_grad()
        close(loss.item(), rl.item(), 2e-4)
    sd = ref.state_dict()
    for name, p in pl.named_parameters():
        close(p.numpy(), sd[mapping[name]].numpy(), 2e-3)
    ev = model.eval_batch([ids[:, :-1], ids[:, 1:]], compute_loss=True)
    with paddle.no_grad():
        ref.eval()
        close(ev.item(), ref(ids[:, :-1], ids[:, 1:]).item(), 2e-3)


def case_pp_bf16_zero_bubble():
    """GPU: bf16 Llama slice on pp=2 with the zero-bubble schedule: activations travel through the peer-memory mailbox, weight gradients
    are parked by B and accumulated into the flat gradient arena by W (tcgen05 accumulate epilogue); loss tracks a single-GPU run."""
    from paddle_b200.kernels import wgrad as WG
    from paddle_b200.models import llama as LL

    s, hcg = setup(pp=2)
    mode = os.environ.get("B200_TEST_PP_MODE", "ZBH1")
    s.pipeline_configs = {"accumulate_steps": 4, "micro_batch_size": 2, "schedule_mode": mode}
    paddle.set_default_dtype("bfloat16")
    cfg = LL.llama_tiny(hidden_size=512, intermediate_size=1024, num_attention_heads=4, num_key_value_heads=4, num_hidden_layers=4,
                        vocab_size=1024, max_position_embeddings=256)
    ids = paddle.to_tensor(np.random.RandomState(0).randint(0, cfg.vocab_size, (8, 257)))
    paddle.seed(11)
    ref = LL.LlamaForCausalLM(cfg)
    ref_sd = {k: v.clone() for k, v in ref.state_dict().items()}
    from paddle_b200.distributed.fleet.pipeline import PipelineLayer

    pl = PipelineLayer(LL.pipeline_layer_descs(cfg), num_stages=2, loss_fn=LL.LlamaPretrainingCriterion(cfg), seg_method="layer:LlamaDecoderLayer")
    mapping = {}
    for name, _ in pl.named_parameters():
        idx, rest = name.split(".", 1)
        i = int(idx)
        mapping[name] = ("llama.embedding." if i == 0 else "lm_head." if i == cfg.num_hidden_layers + 1 else f"llama.layers.{i - 1}.") + rest
    pl.set_state_dict({k: ref_sd[v] for k, v in mapping.items()})
    model = fleet.distributed_model(pl)
    mk = lambda ps: paddle.optimizer.AdamW(1e-3, parameters=ps, weight_decay=0.0, multi_precision=True, grad_clip=paddle.nn.ClipGradByGlobalNorm(1.0))  # noqa: E731
    opt = fleet.distributed_optimizer(mk(pl.parameters()))
    ropt = mk(ref.parameters())
    ropt.enable_flat_arena()
    WG.stats.update(parked=0, fused=0, returned=0)
    losses, rlosses = [], []
    for _ in range(4):
        losses.append(float(model.train_batch([ids[:, :-1], ids[:, 1:]], opt).item()))
        acc = 0.0
        for mb in range(4):
            sl = slice(2 * mb, 2 * mb + 2)
            l = ref(ids[sl, :-1], ids[sl, 1:]) / 4
            l.backward()
            acc += float(l.item())
        ropt.step()
        ropt.clear_grad()
        rlosses.append(acc)
    if GPU:
        assert model.transport_name() == "mailbox", model.transport_name()
        w = model.exposed_wait()
        assert w is not None and w["waits"] > 0, w
        assert WG.stats["fused"] > 0, WG.stats                      # the single-GPU reference accumulates straight into its arena
        if mode.upper().startswith("ZB"):
            assert WG.stats["parked"] > 0, WG.stats                 # B passes parked their weight gradients for W
    for a, b in zip(losses, rlosses):
        assert abs(a - b) < 0.05 * abs(b) + 0.02, (losses, rlosses)
    assert losses[-1] < losses[0]
    paddle.set_default_dtype("float32")


def case_pp_interleave():
    """Virtual pipeline (pp=2 x 2 chunks per rank, 4 micro-batches) == single-process training. Parity: hybrid_parallel_pp_interleave*.py."""
    s, hcg = setup(pp=2)
    L, cfg = _tiny_llama({"num_hidden_layers": 4})
    ids = paddle.to_tensor(np.random.RandomState(0).randint(0, cfg.vocab_size, (4, 33)))
    s.pipeline_configs = {"accumulate_steps": 4, "micro_batch_size": 1, "schedule_mode": os.environ.get("B200_TEST_PP_MODE", "1F1B")}
    paddle.seed(11)
    from paddle_b200.distributed.fleet.pipeline import PipelineLayer, PipelineParallelWithInterleave

    ref, _ = _train_ref(L, cfg, ids, 0)
    ref_sd = {k: v.clone() for k, v in ref.state_dict().items()}
    pl = PipelineLayer(L.pipeline_layer_descs(cfg), num_stages=2, loss_fn=L.LlamaPretrainingCriterion(cfg), seg_method="layer:LlamaDecoderLayer",
                       num_virtual_pipeline_stages=2)
    mapping = {}
    for name, _ in pl.named_parameters():
        idx, rest = name.split(".", 1)
        i = int(idx.split("_")[-1])
        if i == 0:
            mapping[name] = "llama.embedding." + rest
        elif i == cfg.num_hidden_layers + 1:
            mapping[name] = "lm_head." + rest
        else:
            mapping[name] = f"llama.layers.{i - 1}." + rest
    pl.set_state_dict({k: ref_sd[v] for k, v in mapping.items()})
    model = fleet.distributed_model(pl)
    assert isinstance(model, PipelineParallelWithInterleave)
    opt = fleet.distributed_optimizer(paddle.optimizer.AdamW(1e-2, parameters=pl.parameters(), weight_decay=0.0))
    ropt = paddle.optimizer.AdamW(1e-2, parameters=ref.parameters(), weight_decay=0.0)
    for _ in range(3):
        loss = model.train_batch([ids[:, :-1], ids[:, 1:]], opt)
        rl = ref(ids[:, :-1], ids[:, 1:])
        rl.backward()
        ropt.step()
        ropt.clear_grad()
        close(loss.item(), rl.item(), 2e-4)
    sd = ref.state_dict()
    for name, p in pl.named_parameters():
        close(p.numpy(), sd[mapping[name]].numpy(), 8e-3)   # Adam turns accumulation-order noise into +-lr steps on near-zero grads


def case_hybrid_mp_pp():
    """mp2 x pp2 (4 ranks) with sequence parallel: loss decreases and matches across ranks."""
    s, hcg = setup(mp=2, pp=2)
    L, cfg = _tiny_llama({"sequence_parallel": True})
    from paddle_b200.distributed.fleet.pipeline import PipelineLayer

    s.pipeline_configs = {"accumulate_steps": 2, "micro_batch_size": 1}
    pl = PipelineLayer(L.pipeline_layer_descs(cfg), num_stages=2, loss_fn=L.LlamaPretrainingCriterion(cfg), seg_method="layer:LlamaDecoderLayer")
    model = fleet.distributed_model(pl)
    opt = fleet.distributed_optimizer(paddle.optimizer.AdamW(5e-3, parameters=pl.parameters(), weight_decay=0.0,
                                                             grad_clip=paddle.nn.ClipGradByGlobalNorm(1.0)))
    ids = paddle.to_tensor(np.random.RandomState(0).randint(0, cfg.vocab_size, (2, 33)))
    losses = [model.train_batch([ids[:, :-1], ids[:, 1:]], opt).item() for _ in range(6)]
    assert losses[-1] < losses[0] - 0.2, losses
    t = paddle.to_tensor([losses[-1]])
    lst = []
    dist.all_gather(lst, t)
    close([x.item() for x in lst], [losses[-1]] * len(lst), 1e-5)


def case_mp_sp_parity():
    """mp2 + sequence parallel Llama (flat-arena AdamW + hybrid global-norm clip) == dense single-process training."""
    _, hcg = setup(mp=2)
    r = hcg.get_model_parallel_rank()
    from paddle_b200.models import llama as L

    cfg_d = L.llama_tiny(dtype="float32")
    ids = paddle.to_tensor(np.random.RandomState(0).randint(0, cfg_d.vocab_size, (2, 33)))
    # dense reference is built with the mp topology hidden, so it uses plain Linear layers
    from paddle_b200.distributed.fleet import topology as topo

    saved = topo.get_hybrid_communicate_group()
    topo._set_hcg(None)
    paddle.seed(11)
    ref = L.LlamaForCausalLM(cfg_d)
    topo._set_hcg(saved)
    for p_ in ref.parameters():  # the lm head draws from the per-rank mp RNG tracker: make the reference identical everywhere
        dist.broadcast(p_, 0)
    cfg = L.llama_tiny(dtype="float32", sequence_parallel=True)
    par = L.LlamaForCausalLM(cfg)
    sd = ref.state_dict()
    h, f = cfg.hidden_size, cfg.intermediate_size

    def shard_cols(w, parts):  # w [in, sum(parts)] -> this rank's slice of every part, concatenated
        outs, off = [], 0
        for n in parts:
            seg = w[:, off:off + n]
            outs.append(seg[:, r * (n // 2):(r + 1) * (n // 2)])
            off += n
        return paddle.concat(outs, axis=1)

    new = {}
    for k, v in par.state_dict().items():
        d = sd[k]
        if k.endswith("qkv_proj.weight"):
            new[k] = shard_cols(d, [h, h, h])
        elif k.endswith("gate_up_proj.weight"):
            new[k] = shard_cols(d, [f, f])
        elif k.endswith("o_proj.weight") or k.endswith("down_proj.weight"):
            n = d.shape[0]
            new[k] = d[r * (n // 2):(r + 1) * (n // 2)]
        elif k.endswith("embed_tokens.weight"):
            n = d.shape[0]
            new[k] = d[r * (n // 2):(r + 1) * (n // 2)]
        elif k == "lm_head.weight":
            n = d.shape[1]
            new[k] = d[:, r * (n // 2):(r + 1) * (n // 2)]
        else:
            new[k] = d
    par.set_state_dict(new)
    model = fleet.distributed_model(par)
    mk = lambda ps: paddle.optimizer.AdamW(1e-2, parameters=ps, weight_decay=0.01, grad_clip=paddle.nn.ClipGradByGlobalNorm(0.5))  # noqa: E731
    opt = fleet.distributed_optimizer(mk(par.parameters()))
    ropt = mk(ref.parameters())
    for _ in range(3):
        loss = model(ids[:, :-1], ids[:, 1:])
        loss.backward()
        opt.step()
        opt.clear_grad()
        rl = ref(ids[:, :-1], ids[:, 1:])
        rl.backward()
        ropt.step()
        ropt.clear_grad()
        print('STEP', _, loss.item(), rl.item(), flush=True)
        close(loss.item(), rl.item(), 5e-4)
    close(par.state_dict()["llama.layers.0.input_layernorm.weight"].numpy(), ref.state_dict()["llama.layers.0.input_layernorm.weight"].numpy(), 2e-3)
    n = ref.state_dict()["llama.layers.1.mlp.down_proj.weight"].shape[0]
    close(par.state_dict()["llama.layers.1.mlp.down_proj.weight"].numpy(),
          ref.state_dict()["llama.layers.1.mlp.down_proj.weight"][r * (n // 2):(r + 1) * (n // 2)].numpy(), 5e-3)


def case_sharding():
    """group_sharded_parallel os_g and p_g_os == plain training (AdamW + global-norm clip + weight decay, gradient accumulation over two
    backward passes, loss scaling with an overflow on one rank), stage-2 optimizer checkpoints are rank-independent.
    Parity: dygraph_group_sharded_stage2/3.py."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    from paddle_b200.distributed.sharding import GroupShardedOptimizerStage2, group_sharded_parallel

    dev_kw = {}
    if GPU:
        paddle.set_device(f"gpu:{os.environ['LOCAL_RANK']}")
    os.environ["B200_SHARD_BUCKET_MB"] = "0.0005"     # several buckets even for this tiny model
    for level in ("os_g", "p_g_os"):
        for kind in ("adamw", "momentum"):
            paddle.seed(21)
            net = nn.Sequential(nn.Linear(8, 32), nn.GELU(), nn.Linear(32, 16), nn.GELU(), nn.Linear(16, 4))
            ref = nn.Sequential(nn.Linear(8, 32), nn.GELU(), nn.Linear(32, 16), nn.GELU(), nn.Linear(16, 4))
            ref.set_state_dict(net.state_dict())
            x, y = paddle.randn([4 * w, 8]), paddle.randn([4 * w, 4])

            def mk(ps):
                clip = paddle.nn.ClipGradByGlobalNorm(0.5)
                if kind == "adamw":
                    return paddle.optimizer.AdamW(1e-2, parameters=ps, weight_decay=0.05, grad_clip=clip)
                return paddle.optimizer.Momentum(0.05, momentum=0.9, parameters=ps, grad_clip=clip)

            opt, ropt = mk(net.parameters()), mk(ref.parameters())
            model, opt, _ = group_sharded_parallel(net, opt, level)
            for it in range(3):
                for half in range(2):                      # two backward passes per step (gradient accumulation)
                    sl = slice(r * 4 + half * 2, r * 4 + half * 2 + 2)
                    loss = ((model(x[sl]) - y[sl]) ** 2).mean() / 2
                    loss.backward()
                opt.step()
                opt.clear_grad()
                rl = ((ref(x) - y) ** 2).mean()
                rl.backward()
                ropt.step()
                ropt.clear_grad()
            sd = model.state_dict()
            for (k, a), (_, b) in zip(sd.items(), ref.state_dict().items()):
                close(a.numpy(), b.numpy(), 2e-3)
            if level == "os_g" and kind == "adamw":
                assert isinstance(opt, GroupShardedOptimizerStage2) and len(opt.arena.slabs[0]["buckets"]) > 1
                osd, rsd = opt.state_dict(), ropt.state_dict()
                for p, rp in zip(net.parameters(), ref.parameters()):
                    assert list(osd[f"{p.name}_moment1_0"].shape) == list(p.shape)
                    close(osd[f"{p.name}_moment1_0"].numpy(), rsd[f"{rp.name}_moment1_0"].numpy(), 2e-3)
                before = {k: v.numpy().copy() for k, v in model.state_dict().items()}
                opt.set_state_dict(osd)                   # round trip keeps training identical
                ((model(x[r * 4:(r + 1) * 4]) - y[r * 4:(r + 1) * 4]) ** 2).mean().backward()
                opt.step()
                opt.clear_grad()
                ((ref(x) - y) ** 2).mean().backward()
                ropt.step()
                ropt.clear_grad()
                for (k, a), (_, b) in zip(model.state_dict().items(), ref.state_dict().items()):
                    close(a.numpy(), b.numpy(), 3e-3)
        # loss scaling: an overflow seen by one rank only skips the update everywhere and halves the scale
        paddle.seed(22)
        net = nn.Sequential(nn.Linear(8, 16), nn.Linear(16, 4))
        opt = paddle.optimizer.AdamW(1e-2, parameters=net.parameters())
        scaler = paddle.amp.GradScaler(init_loss_scaling=1024.0, decr_every_n_nan_or_inf=1, incr_every_n_steps=1000)
        model, opt, scaler = group_sharded_parallel(net, opt, level, scaler=scaler)
        x = paddle.randn([4, 8])
        w0 = {k: v.numpy().copy() for k, v in model.state_dict().items()}
        for it in range(2):
            xin = x * (float("inf") if (it == 0 and r == 1) else 1.0)
            loss = (model(xin) ** 2).mean()
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            opt.clear_grad()
            now = {k: v.numpy().copy() for k, v in model.state_dict().items()}
            if it == 0:
                for k in w0:
                    close(now[k], w0[k], 1e-7)
                assert abs(float(scaler._scale) - 512.0) < 1e-3
            else:
                assert any(not np.allclose(now[k], w0[k]) for k in w0)
    os.environ.pop("B200_SHARD_BUCKET_MB", None)


def case_auto_parallel():
    """DistTensor: shard/reshard round trips, matmul/elementwise/reduce propagation and dp+mp training parity.
    Parity: test/auto_parallel/{test_shard_tensor_api,reshard_*,semi_auto_parallel_simple_net}.py."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    import paddle_b200.distributed as D
    from paddle_b200.distributed.auto_parallel import Partial, ProcessMesh, Replicate, Shard

    mesh = ProcessMesh(list(range(w)), dim_names=["x"])
    paddle.seed(5)
    g = paddle.randn([4 * w, 6 * w])
    a = D.shard_tensor(g, mesh, [Shard(0)])
    assert a.shape == [4 * w, 6 * w] and a._local_shape == [4, 6 * w]
    close(a._local_value().numpy(), g[r * 4:(r + 1) * 4].numpy())
    b = D.reshard(a, mesh, [Shard(1)])
    close(b._local_value().numpy(), g[:, r * 6:(r + 1) * 6].numpy())
    c = D.reshard(b, mesh, [Replicate()])
    close(c._local_value().numpy(), g.numpy())
    p = D.reshard(c, mesh, [Partial()])
    close(D.reshard(p, mesh, [Shard(0)])._local_value().numpy(), g[r * 4:(r + 1) * 4].numpy())
    close(D.unshard_dtensor(a).numpy(), g.numpy())
    # elementwise with a replicated bias that must be sliced, and reductions
    bias = paddle.randn([6 * w])
    e = b + bias
    close(D.unshard_dtensor(e).numpy(), (g + bias).numpy())
    close(float(D.unshard_dtensor(a.sum())), float(g.sum()), 1e-3)
    close(D.unshard_dtensor(b.mean(0)).numpy(), g.mean(0).numpy(), 1e-5)
    close(D.unshard_dtensor(paddle.transpose(a, [1, 0])).numpy(), g.t().numpy())

    # tensor-parallel MLP through propagation rules: col-parallel then row-parallel, fwd + grads vs dense
    paddle.seed(9)
    w1, w2, x = paddle.randn([8, 16]) * 0.3, paddle.randn([16, 8]) * 0.3, paddle.randn([4, 8])
    rw1, rw2 = paddle.to_tensor(w1.numpy(), stop_gradient=False), paddle.to_tensor(w2.numpy(), stop_gradient=False)
    ref = paddle.matmul(paddle.nn.functional.gelu(paddle.matmul(x, rw1)), rw2)
    ref.sum().backward()
    d1 = D.shard_tensor(w1, mesh, [Shard(1)], stop_gradient=False)
    d2 = D.shard_tensor(w2, mesh, [Shard(0)], stop_gradient=False)
    h = paddle.nn.functional.gelu(paddle.matmul(D.shard_tensor(x, mesh, [Replicate()]), d1))
    assert h.placements == [Shard(1)], h.placements
    out = paddle.matmul(h, d2)
    assert out.placements == [Partial()], out.placements
    full = D.reshard(out, mesh, [Replicate()])
    close(full._local_value().numpy(), ref.numpy(), 1e-4)
    full.sum().backward()
    close(d1.grad._local_value().numpy(), rw1.grad[:, r * (16 // w):(r + 1) * (16 // w)].numpy(), 1e-4)
    close(d2.grad._local_value().numpy(), rw2.grad[r * (16 // w):(r + 1) * (16 // w)].numpy(), 1e-4)

    # data parallel: batch sharded, params replicated, shard_optimizer (ZeRO-1 ownership) == single-process training
    paddle.seed(13)
    net = nn.Sequential(nn.Linear(8, 16), nn.Tanh(), nn.Linear(16, 2))
    refn = nn.Sequential(nn.Linear(8, 16), nn.Tanh(), nn.Linear(16, 2))
    refn.set_state_dict(net.state_dict())
    D.shard_layer(net, mesh)
    opt = D.shard_optimizer(paddle.optimizer.AdamW(1e-2, parameters=net.parameters()), D.ShardingStage1("x", mesh))
    ropt = paddle.optimizer.AdamW(1e-2, parameters=refn.parameters())
    X, Y = paddle.randn([4 * w, 8]), paddle.randn([4 * w, 2])
    for _ in range(3):
        xs, ys = D.shard_tensor(X, mesh, [Shard(0)]), D.shard_tensor(Y, mesh, [Shard(0)])
        loss = paddle.nn.functional.mse_loss(net(xs), ys)
        loss.backward()
        opt.step()
        opt.clear_grad()
        rl = paddle.nn.functional.mse_loss(refn(X), Y)
        rl.backward()
        ropt.step()
        ropt.clear_grad()
        close(float(D.unshard_dtensor(loss)), float(rl), 1e-4)
    for (k, pa), (_, pb) in zip(net.state_dict().items(), refn.state_dict().items()):
        close(D.unshard_dtensor(pa).numpy(), pb.numpy(), 1e-4)

    # intermediate API: parallelize() with a col/row plan == dense
    paddle.seed(17)
    mesh2 = ProcessMesh(np.arange(w).reshape(1, w), dim_names=["dp", "mp"])

    class MLP(nn.Layer):
        def __init__(self):
            super().__init__()
            self.up, self.down = nn.Linear(8, 16), nn.Linear(16, 8)

        def forward(self, t):
            return self.down(paddle.nn.functional.relu(self.up(t)))

    m, mr = MLP(), MLP()
    mr.set_state_dict(m.state_dict())
    xin = paddle.randn([4, 8])
    m = D.parallelize(m, mesh=mesh2, config={"mp_config": {"parallelize_plan": {"up": D.ColWiseParallel(), "down": D.RowWiseParallel()}}})
    close(m(xin).numpy(), mr(xin).numpy(), 1e-4)


def case_spmd_rules():
    """SPMD rule library (auto_parallel/spmd_rules.py): ops on sharded DistTensors give the dense result and keep the sharding whenever
    the op does not touch the sharded dimension. Parity: test/auto_parallel/spmd_rules/test_*_rule.py."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    import paddle_b200.distributed as D
    from paddle_b200.distributed.auto_parallel import ProcessMesh, Replicate, Shard

    mesh = ProcessMesh(list(range(w)), dim_names=["x"])
    paddle.seed(5)
    g = paddle.randn([4 * w, 6, 8])
    a = D.shard_tensor(g, mesh, [Shard(0)])
    gt = g.as_subclass(torch.Tensor)

    def check(dt, dense, placements=None, tol=1e-5):
        full = D.unshard_dtensor(dt) if hasattr(dt, "placements") else dt
        close(torch.as_tensor(full.numpy() if hasattr(full, "numpy") else full).double().numpy(), dense.double().numpy(), tol)
        if placements is not None:
            assert list(dt.placements) == placements, (dt.placements, placements)

    check(torch.cumsum(a, dim=1), torch.cumsum(gt, 1), [Shard(0)])
    check(torch.cumsum(a, dim=0), torch.cumsum(gt, 0))                       # touches the sharded dim: un-sharded just for this op
    check(torch.argmax(a, dim=2), gt.argmax(2), [Shard(0)])
    check(torch.amax(a, dim=1), gt.amax(1), [Shard(0)])
    check(paddle.argmax(a, axis=1), gt.argmax(1), [Shard(0)])
    check(torch.unsqueeze(a, 0), gt.unsqueeze(0), [Shard(1)])
    check(paddle.unsqueeze(a, -1), gt.unsqueeze(-1), [Shard(0)])
    check(torch.squeeze(torch.narrow(a, 1, 0, 1), 1), gt[:, :1].squeeze(1), [Shard(0)])
    check(torch.flatten(a, 1, 2), gt.flatten(1, 2), [Shard(0)])
    check(paddle.reshape(a, [4 * w, 48]), gt.reshape(4 * w, 48), [Shard(0)])
    check(paddle.reshape(a, [4 * w * 6, 8]), gt.reshape(4 * w * 6, 8))
    check(torch.cat([a, a], dim=1), torch.cat([gt, gt], 1), [Shard(0)])
    check(torch.stack([a, a], dim=0), torch.stack([gt, gt], 0), [Shard(1)])
    parts = torch.split(a, 3, dim=1)
    check(parts[1], torch.split(gt, 3, 1)[1], [Shard(0)])
    check(torch.flip(a, dims=[2]), torch.flip(gt, [2]), [Shard(0)])
    check(torch.triu(a), torch.triu(gt), [Shard(0)])
    check(torch.topk(a, 2, dim=2)[0], torch.topk(gt, 2, 2)[0], [Shard(0)])
    check(paddle.expand(a, [2, 4 * w, 6, 8]), gt.expand(2, 4 * w, 6, 8), [Shard(1)])
    check(torch.nn.functional.pad(a, (1, 1)), torch.nn.functional.pad(gt, (1, 1)), [Shard(0)])
    idx = paddle.to_tensor(np.array([0, 2, 5]))
    check(torch.index_select(a, 1, idx.as_subclass(torch.Tensor)), torch.index_select(gt, 1, idx.as_subclass(torch.Tensor)), [Shard(0)])
    q = D.shard_tensor(paddle.randn([2 * w, 2, 5, 4]), mesh, [Shard(0)])
    qt = D.unshard_dtensor(q).as_subclass(torch.Tensor)
    check(torch.nn.functional.scaled_dot_product_attention(q, q, q), torch.nn.functional.scaled_dot_product_attention(qt, qt, qt), [Shard(0)], 1e-4)
    img = D.shard_tensor(paddle.randn([2 * w, 3, 8, 8]), mesh, [Shard(0)])
    wt = paddle.randn([4, 3, 3, 3])
    it = D.unshard_dtensor(img).as_subclass(torch.Tensor)
    check(torch.nn.functional.conv2d(img, wt.as_subclass(torch.Tensor)), torch.nn.functional.conv2d(it, wt.as_subclass(torch.Tensor)), [Shard(0)], 1e-4)
    # gradients flow through a rule-driven op chain like through the dense one
    gp = paddle.to_tensor(g.numpy(), stop_gradient=False)
    ap = D.shard_tensor(paddle.to_tensor(g.numpy()), mesh, [Shard(0)], stop_gradient=False)
    (torch.flatten(torch.cumsum(ap, dim=2), 1, 2) ** 2).sum().backward()
    (torch.cumsum(gp.as_subclass(torch.Tensor), 2).flatten(1, 2) ** 2).sum().backward()
    close(ap.grad._local_value().numpy(), gp.grad[r * 4:(r + 1) * 4].numpy(), 1e-4)


def case_dist_checkpoint():
    """Sharded save on one layout, load on another. Parity: test/auto_parallel/semi_auto_parallel_checkpoint_*.py."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    import paddle_b200.distributed as D
    from paddle_b200.distributed.auto_parallel import ProcessMesh, Replicate, Shard

    mesh = ProcessMesh(list(range(w)), dim_names=["x"])
    paddle.seed(3)
    g1, g2 = paddle.randn([4 * w, 6 * w]), paddle.randn([5])
    path = os.environ.get("B200_TEST_TMP", "/tmp") + "/dist_ckpt_case"
    sd = {"a": D.shard_tensor(g1, mesh, [Shard(0)]), "b": D.shard_tensor(g2, mesh, [Replicate()])}
    D.save_state_dict(sd, path)
    tgt = {"a": D.shard_tensor(paddle.zeros([4 * w, 6 * w]), mesh, [Shard(1)]), "b": D.shard_tensor(paddle.zeros([5]), mesh, [Replicate()])}
    D.load_state_dict(tgt, path)
    close(tgt["a"]._local_value().numpy(), g1[:, r * 6:(r + 1) * 6].numpy())
    close(tgt["b"]._local_value().numpy(), g2.numpy())


def case_dist_checkpoint_mp():
    """fleet tensor-parallel layers through the distributed checkpoint: every mp rank must get ITS OWN block back (shard metadata on
    Column / Row / VocabParallel parameters and on their optimizer state), not the last writer's."""
    s, hcg = setup(mp=2)
    r = hcg.get_model_parallel_rank()
    import paddle_b200.distributed as D
    from paddle_b200.distributed.fleet import mp_layers as mpu

    paddle.seed(7)
    net = nn.Sequential(mpu.VocabParallelEmbedding(16, 8), mpu.ColumnParallelLinear(8, 12, has_bias=True, gather_output=False),
                        mpu.RowParallelLinear(12, 8, has_bias=True, input_is_parallel=True))
    fleet.distributed_model(net)
    opt = paddle.optimizer.AdamW(1e-2, parameters=net.parameters())
    ids = paddle.to_tensor(np.arange(8).reshape(2, 4))
    (net(ids) ** 2).mean().backward()
    opt.step()
    opt.clear_grad()
    assert net[1].weight.__dict__["_dist_shard"] == ((8, 12), (0, 6 * r)) and net[2].weight.__dict__["_dist_shard"][1] == (6 * r, 0)
    want = {k: v.numpy().copy() for k, v in net.state_dict().items()}
    osd = opt.state_dict()
    want_m = {k: v.numpy().copy() for k, v in osd.items() if k.endswith("_moment1_0")}
    path = os.environ.get("B200_TEST_TMP", "/tmp") + "/dist_ckpt_mp_case"
    D.save_state_dict(net.state_dict(), path + "/model")
    D.save_state_dict({k: v for k, v in osd.items() if hasattr(v, "shape")}, path + "/opt")
    with paddle.no_grad():
        for p in net.parameters():
            p.zero_()
    sd = net.state_dict()
    D.load_state_dict(sd, path + "/model")
    for k, v in net.state_dict().items():
        close(v.numpy(), want[k], 1e-6)
    tgt = {k: v for k, v in opt.state_dict().items() if hasattr(v, "shape")}
    for k, v in tgt.items():
        if k.endswith("_moment1_0"):
            with paddle.no_grad():
                v.zero_()
    D.load_state_dict(tgt, path + "/opt")
    for k, v in want_m.items():
        close(tgt[k].numpy(), v, 1e-6)
    # the two mp ranks really hold different blocks (otherwise the test proves nothing)
    w = torch.tensor(want["1.weight"])
    both = [torch.zeros_like(w) for _ in range(2)]
    torch.distributed.all_gather(both, w)
    assert not torch.allclose(both[0], both[1])


def case_p2p_kernels():
    """Peer-memory collectives vs NCCL (GPU only)."""
    assert GPU
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    from paddle_b200.parallel import symm

    sc = symm.context_for(None)
    assert sc is not None, "symmetric heap unavailable"
    torch.manual_seed(r)
    for n in (1 << 10, (1 << 20) + 8, 3 << 20):
        for dt in (torch.bfloat16, torch.float32):
            t = torch.randn(n, device="cuda").to(dt)
            ref = t.clone()
            torch.distributed.all_reduce(ref)
            sc.allreduce_(t)
            torch.cuda.synchronize()
            assert (t.float() - ref.float()).abs().max().item() <= 2e-2 * max(1.0, ref.float().abs().max().item()), (n, dt)
    # fused row-parallel linear etc.
    from paddle_b200.parallel import fused_mp

    x = (torch.randn(256, 512, device="cuda") * 0.1).to(torch.bfloat16)
    wgt = (torch.randn(512, 384, device="cuda") * 0.1).to(torch.bfloat16)
    grp = dist.collective._global_group()
    y = fused_mp.row_parallel_linear(x, wgt, grp)
    ref = (x.float() @ wgt.float())
    torch.distributed.all_reduce(ref)
    assert ((y.float() - ref).norm() / ref.norm()).item() < 2e-2
    # GEMM -> reduce-scatter (sequence dim) and all-gather -> GEMM, forward and backward, vs NCCL + fp32 matmul
    torch.manual_seed(100 + r)
    xs = (torch.randn(128 * w, 2, 512, device="cuda") * 0.1).to(torch.bfloat16).requires_grad_(True)      # [S, B, K_local]
    wr = (torch.randn(512, 384, device="cuda") * 0.1).to(torch.bfloat16).requires_grad_(True)
    yrs = fused_mp.linear_reduce_scatter(xs, wr, grp)
    full = xs.detach().float() @ wr.detach().float()
    torch.distributed.all_reduce(full)
    ref_rs = full.chunk(w, 0)[r]
    assert list(yrs.shape) == list(ref_rs.shape), (yrs.shape, ref_rs.shape)
    e = ((yrs.float() - ref_rs).norm() / ref_rs.norm()).item()
    assert e < 2e-2, f"gemm_reduce_scatter fwd {e}"
    gy = (torch.randn_like(yrs) * 0.1)
    yrs.backward(gy)
    gl = [torch.empty_like(gy) for _ in range(w)]
    torch.distributed.all_gather(gl, gy.contiguous())
    gfull = torch.cat(gl, 0).float()
    ref_dx = gfull @ wr.detach().float().t()
    ref_dw = xs.detach().float().reshape(-1, 512).t() @ gfull.reshape(-1, 384)
    e = ((xs.grad.float() - ref_dx).norm() / ref_dx.norm()).item()
    assert e < 2e-2, f"gemm_reduce_scatter dx {e}"
    e = ((wr.grad.float() - ref_dw).norm() / ref_dw.norm()).item()
    assert e < 2e-2, f"gemm_reduce_scatter dw {e}"
    xa = (torch.randn(128, 2, 512, device="cuda") * 0.1).to(torch.bfloat16).requires_grad_(True)          # [S/p, B, K]
    wc = (torch.randn(512, 256, device="cuda") * 0.1).to(torch.bfloat16).requires_grad_(True)
    yag = fused_mp.allgather_linear(xa, wc, grp)
    xl = [torch.empty_like(xa) for _ in range(w)]
    torch.distributed.all_gather(xl, xa.detach().contiguous())
    xfull = torch.cat(xl, 0).float()
    ref_ag = xfull @ wc.detach().float()
    e = ((yag.float() - ref_ag).norm() / ref_ag.norm()).item()
    assert e < 2e-2, f"allgather_gemm fwd {e}"
    g2 = torch.randn_like(yag) * 0.1
    yag.backward(g2)
    dxf = g2.float() @ wc.detach().float().t()
    torch.distributed.all_reduce(dxf)
    e = ((xa.grad.float() - dxf.chunk(w, 0)[r]).norm() / dxf.chunk(w, 0)[r].norm()).item()
    assert e < 2e-2, f"allgather_gemm dx {e}"
    ref_dw2 = xfull.reshape(-1, 512).t() @ g2.float().reshape(-1, 256)
    e = ((wc.grad.float() - ref_dw2).norm() / ref_dw2.norm()).item()
    assert e < 2e-2, f"allgather_gemm dw {e}"
    # repeated larger calls through the same symmetric buffers (slot / flag reuse, start+end barriers) with no host sync between them
    sc2 = symm.context_for(grp)
    for it in range(4):
        torch.manual_seed(1000 + 10 * it + r)
        xb = (torch.randn(1024, 1024, device="cuda") * 0.05).to(torch.bfloat16)
        wb = (torch.randn(1024, 1536, device="cuda") * 0.05).to(torch.bfloat16)
        ya = sc2.allgather_gemm(xb, wb)
        yr = sc2.gemm_reduce_scatter(ya[:, :1024].contiguous(), wb)
        gl = [torch.empty_like(xb) for _ in range(w)]
        torch.distributed.all_gather(gl, xb)
        ref_a = torch.cat(gl, 0).float() @ wb.float()
        e = ((ya.float() - ref_a).norm() / ref_a.norm()).item()
        assert e < 2e-2, f"iter {it} fused all-gather GEMM {e}"
        ref_r = ya[:, :1024].float() @ wb.float()
        torch.distributed.all_reduce(ref_r)
        ref_r = ref_r.chunk(w, 0)[r]
        e = ((yr.float() - ref_r).norm() / ref_r.norm()).item()
        assert e < 2e-2, f"iter {it} fused GEMM reduce-scatter {e}"
    # vocab-parallel fused cross entropy kernels vs dense fp32
    from paddle_b200.kernels import loss as KL

    torch.manual_seed(7)
    logits = (torch.randn(300, 1024, device="cuda") * 2).to(torch.bfloat16)
    labels = torch.randint(0, 1024, (300,), device="cuda")
    labels[3] = -100
    v = 1024 // w
    loc = logits[:, r * v:(r + 1) * v].contiguous().requires_grad_(True)
    lv = KL.vocab_parallel_cross_entropy(loc, labels, r * v, None, -100)
    lref_in = logits.float().requires_grad_(True)
    lref = torch.nn.functional.cross_entropy(lref_in, labels, ignore_index=-100, reduction="none")
    e = ((lv.float() - lref).norm() / lref.norm()).item()
    assert e < 1e-2, f"vocab parallel CE fwd {e}"
    lv.sum().backward()
    lref.sum().backward()
    gref = lref_in.grad[:, r * v:(r + 1) * v]
    e = ((loc.grad.float() - gref).norm() / gref.norm()).item()
    assert e < 3e-2, f"vocab parallel CE bwd {e}"
    dist.barrier()


def case_stress_dp_mp_overlap():
    """Deadlock scenario of the round-1 review: a spinning peer-memory all-reduce of the DATA-PARALLEL group on a side stream while the main
    stream runs persistent 148-CTA tcgen05 GEMMs that themselves wait on peers of the MODEL-PARALLEL group (fused all-gather -> GEMM and
    GEMM -> reduce-scatter).  Both must make progress on every rank for many iterations (each group has its own heap / signal pad, and
    the all-reduce grid leaves SM room next to the GEMM CTAs); results are checked against NCCL at the end."""
    assert GPU
    s, hcg = setup(dp=2, mp=2)
    from paddle_b200.parallel import fused_mp, symm

    dp_g, mp_g = hcg.get_data_parallel_group(), hcg.get_model_parallel_group()
    sc_dp = symm.context_for(dp_g, heap_bytes=(256 << 20))
    sc_mp = symm.context_for(mp_g)
    assert sc_dp is not None and sc_mp is not None and sc_dp is not sc_mp
    n = 8 << 20
    grad, _ = sc_dp.buffer("stress_grad", (n,), torch.bfloat16)
    side = torch.cuda.Stream()
    torch.manual_seed(hcg.get_model_parallel_rank())
    xa = (torch.randn(512, 1, 1024, device="cuda") * 0.05).to(torch.bfloat16)          # [S/p, B, K]
    wc = (torch.randn(1024, 1024, device="cuda") * 0.05).to(torch.bfloat16)
    xr = (torch.randn(1024, 1, 512, device="cuda") * 0.05).to(torch.bfloat16)         # [S, B, K_local]
    wr = (torch.randn(512, 1024, device="cuda") * 0.05).to(torch.bfloat16)
    iters = int(os.environ.get("B200_STRESS_ITERS", "1000"))
    y1 = y2 = None
    for it in range(iters):
        grad.fill_(1.0)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            sc_dp.allreduce_(grad)                       # spins on the dp peer
        y1 = fused_mp.allgather_linear(xa, wc, mp_g)    # persistent GEMM, copy warps pull the mp peer's shard
        y2 = fused_mp.linear_reduce_scatter(xr, wr, mp_g)
        torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert float(grad.float().min()) == 2.0 and float(grad.float().max()) == 2.0
    xl = [torch.empty_like(xa) for _ in range(2)]
    torch.distributed.all_gather(xl, xa, group=mp_g.pg)
    ref1 = torch.cat(xl, 0).float() @ wc.float()
    assert ((y1.as_subclass(torch.Tensor).float() - ref1).norm() / ref1.norm()).item() < 2e-2
    full = xr.float() @ wr.float()
    torch.distributed.all_reduce(full, group=mp_g.pg)
    ref2 = full.chunk(2, 0)[hcg.get_model_parallel_rank()]
    assert ((y2.as_subclass(torch.Tensor).float() - ref2).norm() / ref2.norm()).item() < 2e-2
    dist.barrier()


def case_moe_fused_a2a():
    """MoE layer over the fused peer-memory dispatch/combine kernel == the NCCL global_scatter/global_gather path (fwd + grads)."""
    assert GPU
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    from paddle_b200.incubate.moe import ExpertFFN, MoELayer

    grp = dist.collective._global_group()

    def build():
        paddle.seed(100 + r)
        return MoELayer(256, ExpertFFN(2, 256, 512, activation="swiglu"), gate={"type": "naive", "top_k": 2}, moe_group=grp)

    torch.manual_seed(7 + r)
    x0 = (torch.randn(4, 96, 256, device="cuda") * 0.5).to(torch.bfloat16)
    outs, grads = [], []
    for fused in (True, False):
        paddle.set_flags({"FLAGS_b200_p2p_collectives": fused})
        paddle.set_default_dtype("bfloat16")
        m = build()
        paddle.set_default_dtype("float32")
        x = x0.clone().as_subclass(paddle.Tensor)
        x.stop_gradient = False
        y = m(x)
        (y.float() ** 2).mean().backward()
        outs.append(y.as_subclass(torch.Tensor).float())
        grads.append((x.grad.as_subclass(torch.Tensor).float(), m.experts.w1.grad.as_subclass(torch.Tensor).float()))
    paddle.set_flags({"FLAGS_b200_p2p_collectives": True})
    e = ((outs[0] - outs[1]).norm() / outs[1].norm()).item()
    assert e < 1e-2, f"fused MoE forward {e}"
    for a, b, name in ((grads[0][0], grads[1][0], "dx"), (grads[0][1], grads[1][1], "dw1")):
        e = ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
        assert e < 2e-2, f"fused MoE {name} {e}"
    dist.barrier()


def case_mp_sp_bf16():
    """bf16 GPU run of the mp2+SP tiny Llama through the fused peer-memory paths: loss tracks the dense bf16 model."""
    assert GPU
    _, hcg = setup(mp=2)
    r = hcg.get_model_parallel_rank()
    from paddle_b200.distributed.fleet import topology as topo
    from paddle_b200.models import llama as L

    paddle.set_default_dtype("bfloat16")
    kw = dict(hidden_size=256, intermediate_size=512, num_attention_heads=4, num_key_value_heads=4, vocab_size=1024, max_position_embeddings=128)
    cfg_d = L.llama_tiny(**kw)
    ids = paddle.to_tensor(np.random.RandomState(0).randint(0, cfg_d.vocab_size, (2, 129)))
    saved = topo.get_hybrid_communicate_group()
    topo._set_hcg(None)
    paddle.seed(11)
    ref = L.LlamaForCausalLM(cfg_d)
    topo._set_hcg(saved)
    for p_ in ref.parameters():
        dist.broadcast(p_, 0)
    cfg = L.llama_tiny(sequence_parallel=True, **kw)
    par = L.LlamaForCausalLM(cfg)
    sd = ref.state_dict()
    h, f = cfg.hidden_size, cfg.intermediate_size

    def shard_cols(wt, parts):
        outs, off = [], 0
        for n in parts:
            seg = wt[:, off:off + n]
            outs.append(seg[:, r * (n // 2):(r + 1) * (n // 2)])
            off += n
        return paddle.concat(outs, axis=1)

    new = {}
    for k, v in par.state_dict().items():
        d = sd[k]
        if k.endswith("qkv_proj.weight"):
            new[k] = shard_cols(d, [h, h, h])
        elif k.endswith("gate_up_proj.weight"):
            new[k] = shard_cols(d, [f, f])
        elif k.endswith("o_proj.weight") or k.endswith("down_proj.weight") or k.endswith("embed_tokens.weight"):
            n = d.shape[0]
            new[k] = d[r * (n // 2):(r + 1) * (n // 2)]
        elif k == "lm_head.weight":
            n = d.shape[1]
            new[k] = d[:, r * (n // 2):(r + 1) * (n // 2)]
        else:
            new[k] = d
    par.set_state_dict(new)
    model = fleet.distributed_model(par)
    mk = lambda ps: paddle.optimizer.AdamW(1e-3, parameters=ps, weight_decay=0.01, multi_precision=True, grad_clip=paddle.nn.ClipGradByGlobalNorm(1.0))  # noqa: E731
    opt = fleet.distributed_optimizer(mk(par.parameters()))
    ropt = mk(ref.parameters())
    for it in range(4):
        loss = model(ids[:, :-1], ids[:, 1:])
        loss.backward()
        opt.step()
        opt.clear_grad()
        rl = ref(ids[:, :-1], ids[:, 1:])
        rl.backward()
        ropt.step()
        ropt.clear_grad()
        print("STEP", it, loss.item(), rl.item(), flush=True)
        assert np.isfinite(loss.item()), "loss is not finite"
        close(loss.item(), rl.item(), 3e-2)
    paddle.set_default_dtype("float32")


def case_moe_ep():
    """Expert parallel MoELayer (2 ranks x 2 local experts) == one process holding all 4 experts (forward, dx, expert grads).
    Parity: test/collective/parallel_dygraph_moe / moe_layer tests."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    from paddle_b200.incubate.moe import MoELayer

    grp = dist.collective._global_group()
    D, EL = 16, 2

    def experts(seed_base, n, first):
        ls = []
        for e in range(n):
            paddle.seed(seed_base + first + e)
            ls.append(nn.Sequential(nn.Linear(D, 32), nn.GELU(), nn.Linear(32, D)))
        return nn.LayerList(ls)

    par_experts, ref_experts = experts(500, EL, r * EL), experts(500, EL * w, 0)
    paddle.seed(3)          # the gate is replicated: same seed on every rank (the experts above reseed per expert)
    par = MoELayer(D, par_experts, gate={"type": "naive", "top_k": 2}, moe_group=grp)
    paddle.seed(3)
    ref = MoELayer(D, ref_experts, gate={"type": "naive", "top_k": 2}, moe_group=None)
    xs = []
    for k in range(w):
        torch.manual_seed(40 + k)
        xs.append(torch.randn(12, D))
    x = xs[r].clone().as_subclass(paddle.Tensor)
    x.stop_gradient = False
    y = par(x)
    (y ** 2).sum().backward()
    ref_x = [t.clone().as_subclass(paddle.Tensor) for t in xs]
    for t in ref_x:
        t.stop_gradient = False
    ref_y = [ref(t) for t in ref_x]
    sum((o ** 2).sum() for o in ref_y).backward()
    close(y.numpy(), ref_y[r].numpy(), 1e-4)
    close(x.grad.numpy(), ref_x[r].grad.numpy(), 1e-4)
    for e in range(EL):
        for (_, pa), (_, pb) in zip(par.experts[e].named_parameters(), ref.experts[r * EL + e].named_parameters()):
            close(pa.grad.numpy(), pb.grad.numpy(), 1e-4)
    # gate weight gradient: each rank sees its own tokens only; the sum over ranks is the single-process gradient
    g = par.gate.gate.weight.grad.clone()
    dist.all_reduce(g)
    close(g.numpy(), ref.gate.gate.weight.grad.numpy(), 1e-4)


def case_dp_no_sync():
    """DataParallel.no_sync gradient accumulation (2 micro-batches, one all-reduce) == one big batch. Parity: parallel_dygraph_no_sync."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    paddle.seed(9)
    net = nn.Sequential(nn.Linear(6, 12), nn.Tanh(), nn.Linear(12, 3))
    ref = nn.Sequential(nn.Linear(6, 12), nn.Tanh(), nn.Linear(12, 3))
    ref.set_state_dict(net.state_dict())
    x, y = paddle.randn([16, 6]), paddle.randn([16, 3])
    dp = paddle.DataParallel(net)
    opt = paddle.optimizer.SGD(0.05, parameters=dp.parameters())
    ropt = paddle.optimizer.SGD(0.05, parameters=ref.parameters())
    for _ in range(2):
        mine = slice(r * 8, (r + 1) * 8)
        xa, ya = x[mine], y[mine]
        with dp.no_sync():
            (((dp(xa[:4]) - ya[:4]) ** 2).mean() / 2).backward()
        (((dp(xa[4:]) - ya[4:]) ** 2).mean() / 2).backward()
        opt.step()
        opt.clear_grad()
        ((ref(x) - y) ** 2).mean().backward()
        ropt.step()
        ropt.clear_grad()
    for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
        close(a.numpy(), b.numpy(), 1e-4)


def case_sep_parallel():
    """sep (segment / context parallel) degree 2: parameters are broadcast over the sep group and gradients of the two sequence halves
    are averaged over the dp x sep group — training on the halves equals training on the full sequence (loss / sep degree). Parity: hybrid_parallel_sep_model.py."""
    s = fleet.DistributedStrategy()
    s.hybrid_configs = {"dp_degree": 1, "mp_degree": 1, "pp_degree": 1, "sep_degree": 2}
    fleet.init(is_collective=True, strategy=s)
    hcg = fleet.get_hybrid_communicate_group()
    r = hcg.get_sep_parallel_rank()
    assert hcg.get_sep_parallel_world_size() == 2
    paddle.seed(11 + r)           # different init per rank: the wrapper must broadcast rank 0's parameters
    net = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Linear(16, 8))
    model = fleet.distributed_model(net)
    paddle.seed(11)
    ref = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Linear(16, 8))
    for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
        close(a.numpy(), b.numpy(), 1e-6)
    opt = fleet.distributed_optimizer(paddle.optimizer.SGD(0.1, parameters=net.parameters()))
    ropt = paddle.optimizer.SGD(0.1, parameters=ref.parameters())
    torch.manual_seed(2)
    x = torch.randn(2, 6, 8).as_subclass(paddle.Tensor)      # [batch, seq, hidden]; position-wise model: sequence halves are independent
    for _ in range(2):
        half = x[:, r * 3:(r + 1) * 3]
        (model(half) ** 2).sum().backward()
        opt.step()
        opt.clear_grad()
        ((ref(x) ** 2).sum() / 2).backward()     # gradients are averaged over the dp x sep group
        ropt.step()
        ropt.clear_grad()
    for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
        close(a.numpy(), b.numpy(), 1e-4)


def case_pp_shared_embedding():
    """Pipeline with tied input / output embedding (SharedLayerDesc on the first and last stage): the shared weight's gradient is
    all-reduced between its two owners, training == single process with a truly tied weight. Parity: hybrid_parallel_shared_weight.py."""
    s, hcg = setup(pp=2)
    from paddle_b200.distributed.fleet.pipeline import LayerDesc, PipelineLayer, SharedLayerDesc

    V, H = 24, 16
    s.pipeline_configs = {"accumulate_steps": 2, "micro_batch_size": 2}

    class Emb(nn.Layer):
        def __init__(self):
            super().__init__()
            self.weight = self.create_parameter([V, H])

        def forward(self, ids):
            return paddle.nn.functional.embedding(ids, self.weight)

    def head(layer, x):
        return paddle.matmul(x, layer.weight, transpose_y=True)

    class Block(nn.Layer):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(H, H)

        def forward(self, x):
            return x + paddle.tanh(self.fc(x))

    def loss_fn(logits, labels):
        return paddle.nn.functional.cross_entropy(logits.reshape([-1, V]), labels.reshape([-1]))

    paddle.seed(31)
    descs = [SharedLayerDesc("embed", Emb), LayerDesc(Block), LayerDesc(Block), SharedLayerDesc("embed", Emb, forward_func=head)]
    pl = PipelineLayer(descs, num_stages=2, loss_fn=loss_fn)
    # single-process reference with the same initial values: gather every stage's parameters
    local = {n: p.numpy() for n, p in pl.named_parameters()}
    allp = [None, None]
    dist.all_gather_object(allp, local)
    paddle.seed(31)
    emb, b0, b1 = Emb(), Block(), Block()
    emb_w = [v for d in allp for n, v in d.items() if "weight" in n and v.shape == (V, H)][0]
    emb.weight.set_value(emb_w)
    blocks = {}
    for d in allp:
        for n, v in d.items():
            if v.shape != (V, H):
                blocks[n] = v
    names = sorted(blocks)      # "1.fc.weight", "1.fc.bias", "2.fc..."
    for blk, idx in ((b0, "1"), (b1, "2")):
        blk.fc.weight.set_value(blocks[f"{idx}.fc.weight"])
        blk.fc.bias.set_value(blocks[f"{idx}.fc.bias"])
    ref_params = list(emb.parameters()) + list(b0.parameters()) + list(b1.parameters())
    model = fleet.distributed_model(pl)
    opt = fleet.distributed_optimizer(paddle.optimizer.SGD(0.1, parameters=pl.parameters()))
    ropt = paddle.optimizer.SGD(0.1, parameters=ref_params)
    rs = np.random.RandomState(1)
    for _ in range(3):
        ids = paddle.to_tensor(rs.randint(0, V, (4, 5)))
        lab = paddle.to_tensor(rs.randint(0, V, (4, 5)))
        loss = model.train_batch([ids, lab], opt)
        rl = loss_fn(head(emb, b1(b0(emb(ids)))), lab)
        rl.backward()
        ropt.step()
        ropt.clear_grad()
        close(loss.item(), rl.item(), 1e-4)
    shared = [p for n, p in pl.named_parameters() if tuple(p.shape) == (V, H)][0]
    close(shared.numpy(), emb.weight.numpy(), 1e-4)


def case_dp_unused_params():
    """DataParallel(find_unused_parameters=True): a branch that is skipped on some steps must not stall the bucket reduction."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()

    class Net(nn.Layer):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.out = nn.Linear(6, 6), nn.Linear(6, 6), nn.Linear(6, 2)

        def forward(self, x, use_b):
            h = self.a(x)
            if use_b:
                h = h + self.b(x)
            return self.out(paddle.tanh(h))

    paddle.seed(13)
    net, ref = Net(), Net()
    ref.set_state_dict(net.state_dict())
    dp = paddle.DataParallel(net, find_unused_parameters=True, comm_buffer_size=1)
    opt = paddle.optimizer.SGD(0.1, parameters=dp.parameters())
    ropt = paddle.optimizer.SGD(0.1, parameters=ref.parameters())
    x, y = paddle.randn([8, 6]), paddle.randn([8, 2])
    for step in range(4):
        use_b = step % 2 == 0
        sl = slice(r * 4, (r + 1) * 4)
        ((dp(x[sl], use_b) - y[sl]) ** 2).mean().backward()
        opt.step()
        opt.clear_grad()
        ((ref(x, use_b) - y) ** 2).mean().backward()
        ropt.step()
        ropt.clear_grad()
    for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
        close(a.numpy(), b.numpy(), 1e-4)


def case_hybrid_scaler():
    """fleet.distributed_scaler under mp=2: an overflow seen by one rank only must skip the update on every rank and shrink the scale."""
    s, hcg = setup(mp=2)
    r = hcg.get_model_parallel_rank()
    paddle.seed(17)
    net = nn.Linear(4, 4)
    model = fleet.distributed_model(net)
    opt = fleet.distributed_optimizer(paddle.optimizer.SGD(0.1, parameters=net.parameters()))
    scaler = fleet.distributed_scaler(paddle.amp.GradScaler(init_loss_scaling=1024.0, decr_every_n_nan_or_inf=1, incr_every_n_steps=1000))
    before = net.weight.numpy().copy()
    x = paddle.ones([2, 4])
    loss = (model(x) ** 2).mean()
    scaler.scale(loss).backward()
    if r == 1:
        net.weight.grad[0, 0] = float("inf")     # overflow on one mp rank only
    scaler.step(opt)
    scaler.update()
    opt.clear_grad()
    close(net.weight.numpy(), before, 1e-7)       # skipped everywhere
    assert abs(float(scaler._scale if hasattr(scaler, "_scale") else scaler.get_loss_scaling()) - 512.0) < 1e-3, "scale must halve on every rank"
    loss = (model(x) ** 2).mean()
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    assert not np.allclose(net.weight.numpy(), before)


def case_hybrid_scaler_dp():
    """fleet.distributed_scaler under dp=2 with the flat-arena AdamW: an overflow on ONE data-parallel replica must not desynchronise the
    gradient all-reduce (it runs before the found_inf decision) and must skip the update on both replicas."""
    s, hcg = setup(dp=2)
    r = hcg.get_data_parallel_rank()
    paddle.seed(23)
    net = nn.Sequential(nn.Linear(4, 8), nn.Linear(8, 4))
    model = fleet.distributed_model(net)
    opt = fleet.distributed_optimizer(paddle.optimizer.AdamW(0.1, parameters=net.parameters()))
    scaler = fleet.distributed_scaler(paddle.amp.GradScaler(init_loss_scaling=1024.0, decr_every_n_nan_or_inf=1, incr_every_n_steps=1000))
    inner = getattr(model, "_layers", model)
    before = [p.numpy().copy() for p in net.parameters()]
    x = paddle.ones([2, 4]) * (r + 1)
    for it in range(3):
        loss = (inner(x) ** 2).mean()
        scaler.scale(loss).backward()
        if it == 0 and r == 1:
            net[0].weight.grad[0, 0] = float("inf")     # overflow on one replica only
        scaler.step(opt)
        scaler.update()
        opt.clear_grad()
        if it == 0:
            for p, b in zip(net.parameters(), before):
                close(p.numpy(), b, 1e-7)                # skipped on both replicas
            assert abs(float(scaler._scale) - 512.0) < 1e-3
    assert not np.allclose(net[0].weight.numpy(), before[0])
    w = torch.tensor(net[0].weight.numpy())
    lst = [torch.zeros_like(w) for _ in range(2)]
    torch.distributed.all_gather(lst, w)
    close(lst[0].numpy(), lst[1].numpy(), 1e-6)          # replicas stayed identical


def case_fleet_sharding_degree():
    """hybrid_configs sharding_degree=2 through fleet.distributed_model / distributed_optimizer == single-process training on the full
    batch, with a global-norm clip that is active (SGD, and AdamW with a large epsilon so that the clip coefficient matters).
    Parity: hybrid_parallel_sharding_model.py."""
    s, hcg = setup(sharding=2)
    r = hcg.get_sharding_parallel_rank()
    for kind in ("sgd", "adamw"):
        def make_opt(params):
            clip = paddle.nn.ClipGradByGlobalNorm(0.05)
            if kind == "sgd":
                return paddle.optimizer.SGD(0.5, parameters=params, grad_clip=clip)
            return paddle.optimizer.AdamW(0.05, parameters=params, weight_decay=0.01, epsilon=1.0, grad_clip=clip)

        paddle.seed(19 + r)
        net = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Linear(16, 4))
        model = fleet.distributed_model(net)
        paddle.seed(19)
        ref = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Linear(16, 4))
        opt = fleet.distributed_optimizer(make_opt(net.parameters()))
        ropt = make_opt(ref.parameters())
        torch.manual_seed(4)
        x, y = torch.randn(8, 8).as_subclass(paddle.Tensor), torch.randn(8, 4).as_subclass(paddle.Tensor)
        for _ in range(3):
            sl = slice(r * 4, (r + 1) * 4)
            ((model(x[sl]) - y[sl]) ** 2).mean().backward()
            opt.step()
            opt.clear_grad()
            ((ref(x) - y) ** 2).mean().backward()
            ropt.step()
            ropt.clear_grad()
        for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
            close(a.numpy(), b.numpy(), 2e-4)
        if kind == "adamw":
            inner = opt._inner_opt
            slabs = inner._arena.all_slabs()
            held, total = sum(sl.state["m"].numel() for sl in slabs), sum(sl.numel for sl in slabs)
            assert held <= total // 2 + 8 * len(slabs), f"stage 1 keeps 1/2 of the moments per rank, holds {held} of {total}"
            sd, rsd = opt.state_dict(), ropt.state_dict()      # checkpoints see full-length state, equal to the unsharded run
            for p, rp in zip(net.parameters(), ref.parameters()):
                close(sd[f"{p.name}_moment1_0"].numpy(), rsd[f"{rp.name}_moment1_0"].numpy(), 2e-4)
                assert list(sd[f"{p.name}_moment2_0"].shape) == list(p.shape)


def case_auto_engine():
    """fleet.auto.Engine.fit over 2 ranks (batch-sharded data, replicated params averaged) == single-process training; evaluate / predict /
    save / load; LocalLayer; to_distributed. Parity: test/auto_parallel/engine_api.py."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    from paddle_b200.distributed.fleet import auto

    class DS(paddle.io.Dataset):
        def __init__(self, n):
            rs = np.random.RandomState(0)
            self.x, self.y = rs.randn(n, 6).astype("float32"), rs.randn(n, 2).astype("float32")

        def __len__(self):
            return len(self.x)

        def __getitem__(self, i):
            return self.x[i], self.y[i]

    paddle.seed(23)
    net = nn.Sequential(nn.Linear(6, 12), nn.Tanh(), nn.Linear(12, 2))
    ref = nn.Sequential(nn.Linear(6, 12), nn.Tanh(), nn.Linear(12, 2))
    ref.set_state_dict(net.state_dict())
    loss = nn.MSELoss()
    eng = auto.Engine(net, loss, paddle.optimizer.SGD(0.1, parameters=net.parameters()), strategy=auto.Strategy())
    ds = DS(16)
    logs = eng.fit(ds, batch_size=8, epochs=2)
    assert len(logs["loss"]) == 4
    ropt = paddle.optimizer.SGD(0.1, parameters=ref.parameters())
    for _ in range(2):
        for lo in (0, 8):
            # DistributedBatchSampler deals sample i to rank i % world: the global batch of a step is samples [lo, lo + 8)
            x, y = paddle.to_tensor(ds.x[lo:lo + 8]), paddle.to_tensor(ds.y[lo:lo + 8])
            loss(ref(x), y).backward()
            ropt.step()
            ropt.clear_grad()
    for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
        close(a.numpy(), b.numpy(), 1e-4)
    ev = eng.evaluate(ds, batch_size=8)
    assert ev["loss"] is not None and ev["loss"] < logs["loss"][0]
    preds = eng.predict(ds, batch_size=8, steps=1)
    assert preds[0].shape == [4, 2]
    tmp = os.environ.get("B200_TEST_TMP", "/tmp")
    eng.save(os.path.join(tmp, "eng", "ck"))
    dist.barrier()
    eng.load(os.path.join(tmp, "eng", "ck"))
    mesh = dist.ProcessMesh([0, 1], dim_names=["x"])

    class Masked(auto.LocalLayer):
        def __init__(self):
            super().__init__(out_dist_attrs=[(mesh, [dist.Partial(dist.ReduceType.kRedSum)])])

        def forward(self, v):
            return (v * (v > 0)).sum().reshape([1])

    full = np.arange(-4, 4, dtype="float32").reshape(4, 2)
    dt = dist.shard_tensor(paddle.to_tensor(full), mesh, [dist.Shard(0)])
    tot = dist.reshard(Masked()(dt), mesh, [dist.Replicate()])
    close(tot.numpy(), [full[full > 0].sum()], 1e-6)
    m2, o2, l2 = dist.to_distributed(net, paddle.optimizer.SGD(0.1, parameters=net.parameters()), paddle.io.DataLoader(ds, batch_size=4), device_num=2)
    xb, yb = next(iter(l2))
    assert list(xb.shape)[0] == 4 * w or list(xb.shape)[0] == 4


def case_recompute_hybrid_partition():
    """recompute_hybrid with the kept inputs sharded over the mp group (and offload flag on): gradients equal the plain forward."""
    s, hcg = setup(mp=2)
    from paddle_b200.distributed.fleet import recompute_hybrid

    paddle.seed(41)
    net = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Dropout(0.2), nn.Linear(16, 8))
    ref = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Dropout(0.2), nn.Linear(16, 8))
    ref.set_state_dict(net.state_dict())
    torch.manual_seed(6)
    x = torch.randn(4, 8)
    res = []
    for m, hybrid in ((net, True), (ref, False)):
        paddle.seed(99)
        inp = x.clone().as_subclass(paddle.Tensor)
        inp.stop_gradient = False
        y = recompute_hybrid({"mp_group": hcg.get_model_parallel_group(), "offload": True, "partition": True}, m, inp) if hybrid else m(inp)
        (y ** 2).sum().backward()
        res.append((y.numpy(), inp.grad.numpy(), [p.grad.numpy() for p in m.parameters()]))
    close(res[0][0], res[1][0], 1e-6)
    close(res[0][1], res[1][1], 1e-5)
    for a, b in zip(res[0][2], res[1][2]):
        close(a, b, 1e-5)


def case_hapi_fit():
    """paddle.Model.fit under 2 trainers (DataParallel + DistributedBatchSampler) == single-process fit on the global batches;
    evaluate() gathers every rank's outputs for the metric."""
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()

    class DS(paddle.io.Dataset):
        def __init__(self, n):
            rs = np.random.RandomState(3)
            self.x = rs.randn(n, 5).astype("float32")
            self.y = (self.x.sum(1) > 0).astype("int64")

        def __len__(self):
            return len(self.x)

        def __getitem__(self, i):
            return self.x[i], self.y[i]

    ds = DS(16)
    paddle.seed(29)
    net = nn.Sequential(nn.Linear(5, 8), nn.ReLU(), nn.Linear(8, 2))
    ref = nn.Sequential(nn.Linear(5, 8), nn.ReLU(), nn.Linear(8, 2))
    ref.set_state_dict(net.state_dict())
    model = paddle.Model(net)
    model.prepare(paddle.optimizer.SGD(0.2, parameters=net.parameters()), nn.CrossEntropyLoss(), paddle.metric.Accuracy())
    model.fit(ds, batch_size=4, epochs=2, shuffle=False, verbose=0)
    ropt = paddle.optimizer.SGD(0.2, parameters=ref.parameters())
    ce = nn.CrossEntropyLoss()
    for _ in range(2):
        for lo in (0, 8):       # per step the two ranks consume samples lo..lo+8 (4 each, dealt round-robin)
            x, y = paddle.to_tensor(ds.x[lo:lo + 8]), paddle.to_tensor(ds.y[lo:lo + 8])
            ce(ref(x), y).backward()
            ropt.step()
            ropt.clear_grad()
    for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
        close(a.numpy(), b.numpy(), 1e-4)
    res = model.evaluate(ds, batch_size=4, verbose=0)
    with paddle.no_grad():
        acc = float((ref(paddle.to_tensor(ds.x)).argmax(-1).numpy() == ds.y).mean())
    close(res["acc"], acc, 1e-6)
    tmp = os.environ.get("B200_TEST_TMP", "/tmp")
    model.save(os.path.join(tmp, "hapi", "ck"))
    dist.barrier()
    assert os.path.exists(os.path.join(tmp, "hapi", "ck.pdparams"))


def case_nvls_kernels():
    """NVSwitch multicast collectives (parallel/nvls.py, csrc/comm/nvls_collectives.cu) vs NCCL.  Exits 77 when the box has no NVLS."""
    assert GPU
    dist.init_parallel_env()
    r, w = dist.get_rank(), dist.get_world_size()
    paddle.set_flags({"FLAGS_b200_nvls": True})
    from paddle_b200.parallel import nvls

    nc = nvls.context_for(None, nbytes=64 << 20)
    if nc is None:
        print("NVLS unavailable on this system")
        sys.exit(77)
    torch.manual_seed(r)
    for n in (1 << 10, 1 << 20, (3 << 20) + 64):
        for dt in (torch.bfloat16, torch.float32, torch.float16):
            t = (torch.randn(n, device="cuda") * 0.5).to(dt)
            ref = t.clone()
            torch.distributed.all_reduce(ref)
            nc.all_reduce_(t)
            torch.cuda.synchronize()
            assert (t.float() - ref.float()).abs().max().item() <= 2e-2 * max(1.0, ref.float().abs().max().item()), (n, dt)
    # a tensor that lives in the buffer is reduced in place (no staging copies)
    v = nc.tensor(1 << 20, (4096, 64), torch.bfloat16)
    v.copy_((torch.randn(4096, 64, device="cuda") * 0.5).to(torch.bfloat16))
    ref = v.clone()
    torch.distributed.all_reduce(ref)
    nc.all_reduce_(v)
    torch.cuda.synchronize()
    assert (v.float() - ref.float()).abs().max().item() <= 2e-2 * max(1.0, ref.float().abs().max().item())
    # reduce-scatter / all-gather
    x = (torch.randn(w * 8192, device="cuda") * 0.5).to(torch.bfloat16)
    full = x.clone()
    torch.distributed.all_reduce(full)
    out = torch.empty(8192, device="cuda", dtype=torch.bfloat16)
    nc.reduce_scatter(out, x)
    torch.cuda.synchronize()
    assert (out.float() - full.chunk(w)[r].float()).abs().max().item() <= 2e-2 * max(1.0, full.float().abs().max().item())
    mine = torch.full((4096,), float(r + 1), device="cuda")
    gathered = torch.empty(w * 4096, device="cuda")
    nc.all_gather(gathered, mine)
    torch.cuda.synchronize()
    assert torch.equal(gathered.view(w, 4096)[:, 0].cpu(), torch.arange(1, w + 1, dtype=torch.float32))
    # the hybrid data-parallel gradient all-reduce takes this path under the flag
    from paddle_b200.distributed.fleet.hybrid import _allreduce_flat

    g = (torch.randn(1 << 18, device="cuda") * 0.1).to(torch.bfloat16)
    ref = g.clone()
    torch.distributed.all_reduce(ref)
    before = nc.epoch
    _allreduce_flat(g, None)
    torch.cuda.synchronize()
    assert nc.epoch == before + 1 and (g.float() - ref.float()).abs().max().item() <= 2e-2


if __name__ == "__main__":
    case = sys.argv[1]
    if GPU:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        paddle.set_device(f"gpu:{os.environ['LOCAL_RANK']}")
    globals()["case_" + case]()
    if dist.is_initialized():
        dist.barrier()
    print(f"rank {os.environ.get('RANK')} case {case} OK")
