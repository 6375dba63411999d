"""BASELINE config 4: GPT-3 6.7B, GroupSharded stage 3 (p_g_os), fp8 linears (FLAGS_b200_fp8_linear), synthetic tokens.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/gpt_stage3_fp8.py"""
import argparse

from common import init_dist, report, timed

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--seq", type=int, default=2048)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--no-fp8", action="store_true")
args = ap.parse_args()
paddle, rank, world = init_dist()
import torch  # noqa: E402

from paddle_b200.distributed.sharding import group_sharded_parallel  # noqa: E402
from paddle_b200.models import gpt as G  # noqa: E402

paddle.seed(0)
paddle.set_default_dtype("bfloat16")
cfg = G.gpt3_6p7b(num_hidden_layers=args.layers, max_position_embeddings=args.seq, recompute=True)
model = G.GPTForCausalLM(cfg)
opt = paddle.optimizer.AdamW(1e-5, parameters=model.parameters(), weight_decay=0.1, multi_precision=True, moment_dtype="bfloat16")
if world > 1:
    model, opt, _ = group_sharded_parallel(model, opt, "p_g_os")
paddle.set_flags({"FLAGS_b200_fp8_linear": not args.no_fp8})
tok = torch.randint(0, cfg.vocab_size, (args.batch, args.seq + 1), device="cuda").as_subclass(paddle.Tensor)


def step():
    loss = model(tok[:, :-1], tok[:, 1:])
    loss.backward()
    opt.step()
    opt.clear_grad()
    return loss


ms, loss = timed(step, args.steps, args.warmup)
report(rank, metric="tokens/sec GPT-3 6.7B GroupSharded stage3 fp8 linears (synthetic)", value=round(args.batch * args.seq * world * args.steps / (ms / 1e3), 1),
       unit="tokens/s", n_gpus=world, ms_per_step=round(ms / args.steps, 2), layers=args.layers, fp8=not args.no_fp8, last_loss=float(loss))
