"""Finite-value audit of one hybrid-parallel Llama training step (run under torchrun). Prints, per step, the loss and the
first non-finite activation / gradient / parameter. Usage:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/debug_nan.py --mp 2 --layers 2"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import paddle_b200 as paddle  # noqa: E402
from paddle_b200.distributed import env, fleet  # noqa: E402
from paddle_b200.models import llama as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mp", type=int, default=2)
ap.add_argument("--layers", type=int, default=2)
ap.add_argument("--seq", type=int, default=4096)
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--no-fused", action="store_true")
ap.add_argument("--accumulate", type=int, default=1)
args = ap.parse_args()

lr_ = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr_)
paddle.set_device(f"gpu:{lr_}")
if args.no_fused:
    os.environ["B200_DISABLE_SYMM"] = "1"
s = fleet.DistributedStrategy()
s.hybrid_configs = {"dp_degree": 1, "mp_degree": args.mp, "pp_degree": 1}
fleet.init(is_collective=True, strategy=s)
rank = env.get_rank()
cfg = L.llama2_13b()
cfg.num_hidden_layers = args.layers
cfg.max_position_embeddings = args.seq
cfg.tensor_parallel_degree = args.mp
cfg.sequence_parallel = args.mp > 1
paddle.seed(1234 + rank)
paddle.set_default_dtype("bfloat16")
model = L.LlamaForCausalLM(cfg)
opt = paddle.optimizer.AdamW(learning_rate=1e-5, beta1=0.9, beta2=0.95, epsilon=1e-8, parameters=model.parameters(), weight_decay=0.1,
                             grad_clip=paddle.nn.ClipGradByGlobalNorm(1.0), multi_precision=True, moment_dtype="bfloat16",
                             apply_decay_param_fun=lambda n: not any(k in n for k in ("norm", "bias")))
names = {id(p): n for n, p in model.named_parameters()}
inner = model
model = fleet.distributed_model(model)
opt = fleet.distributed_optimizer(opt)

bad = []


def hook(name):
    def f(layer, inp, out):
        o = out[0] if isinstance(out, (tuple, list)) else out
        if isinstance(o, torch.Tensor) and o.is_floating_point() and name.count(".") <= 2 and len(arms) < 60:
            arms.append((name, float(o.as_subclass(torch.Tensor).float().pow(2).mean().sqrt())))
        if isinstance(o, torch.Tensor) and o.is_floating_point() and not bad:
            if not bool(torch.isfinite(o.as_subclass(torch.Tensor)).all()):
                bad.append(name)
                print(f"[rank {rank}] NON-FINITE activation after {name} shape={list(o.shape)}", flush=True)
        if name in ("llama", "lm_head.norm") and isinstance(o, torch.Tensor):
            r = o.as_subclass(torch.Tensor).detach().float().reshape(-1, o.shape[-1]).pow(2).mean(-1).sqrt()
            print(f"[rank {rank}] fwd {name}: shape {list(o.shape)} strides {o.stride()} row-rms min {float(r.min()):.4g} (row {int(r.argmin())}) "
                  f"max {float(r.max()):.4g} (row {int(r.argmax())}) median {float(r.median()):.4g}", flush=True)
        if isinstance(o, torch.Tensor) and o.requires_grad:
            def gh(g, name=name):
                gr = g.as_subclass(torch.Tensor)
                if name in ("llama", "lm_head.norm"):
                    rr = gr.detach().float().reshape(-1, gr.shape[-1]).abs().amax(-1)
                    print(f"[rank {rank}] bwd grad into {name}: shape {list(gr.shape)} strides {gr.stride()} row-max: max {float(rr.max()):.4g} (row {int(rr.argmax())}) "
                          f"median {float(rr.median()):.4g} top rows {rr.topk(5).indices.tolist()}", flush=True)
                if len(gbad) < 3 and not bool(torch.isfinite(gr).all()):
                    gbad.append(name)
                    print(f"[rank {rank}] NON-FINITE grad flowing into output of {name} shape={list(g.shape)} "
                          f"nan={int(torch.isnan(gr).sum())} inf={int(torch.isinf(gr).sum())}", flush=True)
                elif len(gbad) == 0 and name.count(".") <= 2:
                    gmax[name] = float(gr.abs().max())
            o.register_hook(gh)
    return f


gbad, gmax, arms = [], {}, []


for n, sub in inner.named_sublayers():
    sub.register_forward_post_hook(hook(n))

gen = torch.Generator().manual_seed(99)   # identical tokens on every mp rank (a replica shares its input)
for step in range(args.steps):
    for mb in range(args.accumulate):
        tok = torch.randint(0, cfg.vocab_size, (args.batch, args.seq + 1), generator=gen).cuda().as_subclass(paddle.Tensor)
        loss = model(tok[:, :-1], tok[:, 1:]) / args.accumulate
        loss.backward()
    if step == 0 and rank == 0:
        print("activation rms in forward order:", [(k, f"{v:.3g}") for k, v in arms[:20]], flush=True)
        print("grad |max| in backward order:", [(k, f"{v:.3g}") for k, v in list(gmax.items())[:14]], flush=True)
    nbad = 0
    for p in inner.parameters():
        g = p.grad
        if g is not None and not bool(torch.isfinite(g.as_subclass(torch.Tensor)).all()) and nbad < 5:
            nbad += 1
            print(f"[rank {rank}] step {step} NON-FINITE grad {names[id(p)]} {list(p.shape)}", flush=True)
    gn = torch.sqrt(sum((p.grad.as_subclass(torch.Tensor).float() ** 2).sum() for p in inner.parameters() if p.grad is not None))
    opt.step()
    opt.clear_grad()
    pbad = [names[id(p)] for p in inner.parameters() if not bool(torch.isfinite(p.as_subclass(torch.Tensor)).all())]
    print(f"[rank {rank}] step {step} loss {float(loss.item()):.4f} local-grad-norm {float(gn):.4f} bad-params {pbad[:4]}", flush=True)
env.destroy_process_group()
