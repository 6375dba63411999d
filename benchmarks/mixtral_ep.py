"""BASELINE config 5: Mixtral-style MoE, experts sharded over all ranks (expert parallel), fused peer-memory dispatch/combine.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/mixtral_ep.py --layers 4"""
import argparse
import os

from common import init_dist, report, timed

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--seq", type=int, default=4096)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--experts", type=int, default=0, help="total experts (default max(8, world))")
ap.add_argument("--nccl", action="store_true", help="use the NCCL global_scatter / global_gather path instead of the fused kernel")
args = ap.parse_args()
paddle, rank, world = init_dist()
import torch  # noqa: E402

from paddle_b200.models import mixtral as M  # noqa: E402

paddle.seed(0)
paddle.set_default_dtype("bfloat16")
group = paddle.distributed.collective._global_group() if world > 1 else None
cfg = M.mixtral_8x7b(num_hidden_layers=args.layers, max_position_embeddings=args.seq, num_local_experts=args.experts or max(8, world))
model = M.MixtralForCausalLM(cfg, moe_group=group)
opt = paddle.optimizer.AdamW(1e-5, parameters=model.parameters(), weight_decay=0.1, multi_precision=True, moment_dtype="bfloat16")
paddle.set_flags({"FLAGS_b200_p2p_collectives": not args.nccl})
tok = torch.randint(0, cfg.vocab_size, (args.batch, args.seq + 1), device="cuda").as_subclass(paddle.Tensor)


def step():
    loss = model(tok[:, :-1], tok[:, 1:])
    loss.backward()
    if world > 1:   # non-expert parameters are data-parallel replicas: one coalesced all-reduce (peer-memory kernel when available)
        from paddle_b200.distributed.fleet.hybrid import _allreduce_tensors

        grads = [p.grad.as_subclass(torch.Tensor) for p in model.parameters() if p.grad is not None and not getattr(p, "no_sync", False)]
        _allreduce_tensors(grads, group, 1.0 / world)
    opt.step()
    opt.clear_grad()
    return loss


if os.environ.get("B200_DEBUG_LOSS"):
    for i in range(6):
        print("rank", rank, "step", i, "loss", float(step()), flush=True)
ms, loss = timed(step, args.steps, args.warmup)
report(rank, metric="tokens/sec Mixtral-style MoE expert parallel (synthetic)", value=round(args.batch * args.seq * world * args.steps / (ms / 1e3), 1),
       unit="tokens/s", n_gpus=world, ms_per_step=round(ms / args.steps, 2), layers=args.layers, fused_a2a=not args.nccl, last_loss=float(loss))
