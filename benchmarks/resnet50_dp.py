"""BASELINE config 2: ResNet-50, paddle.DataParallel, bf16 AMP, synthetic ImageNet batches.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/resnet50_dp.py --batch 256"""
import argparse

from common import init_dist, report, timed

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
paddle, rank, world = init_dist()
import torch  # noqa: E402

paddle.seed(0)
net = paddle.vision.models.resnet50()
net = paddle.DataParallel(net) if world > 1 else net
opt = paddle.optimizer.Momentum(0.1, momentum=0.9, parameters=net.parameters(), weight_decay=1e-4, multi_precision=True)
x = torch.randn(args.batch, 3, 224, 224, device="cuda").as_subclass(paddle.Tensor)
y = torch.randint(0, 1000, (args.batch,), device="cuda").as_subclass(paddle.Tensor)


def step():
    with paddle.amp.auto_cast(level="O1", dtype="bfloat16"):
        loss = paddle.nn.functional.cross_entropy(net(x), y)
    loss.backward()
    opt.step()
    opt.clear_grad()
    return loss


ms, loss = timed(step, args.steps, args.warmup)
report(rank, metric="images/sec ResNet-50 DataParallel bf16 (synthetic)", value=round(args.batch * world * args.steps / (ms / 1e3), 1), unit="images/s",
       n_gpus=world, ms_per_step=round(ms / args.steps, 2), per_gpu_batch=args.batch, last_loss=float(loss))
