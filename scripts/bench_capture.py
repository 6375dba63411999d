"""Eager vs CUDA-graph-captured training step (paddle.jit.capture_train_step) on launch-bound models.
usage: python scripts/bench_capture.py [out.md]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paddle_b200 as paddle  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench_capture.md"
paddle.set_device("gpu:0")


def timeit(fn, n):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def bench(name, make, inputs, n=50):
    if os.environ.get("BENCH_ONLY") and os.environ["BENCH_ONLY"] not in name:
        return None
    rows = []
    for mode in ("eager", "captured"):
        paddle.seed(0)
        net, loss_fn = make()
        opt = paddle.optimizer.AdamW(1e-3, parameters=net.parameters(), weight_decay=0.01, multi_precision=True, grad_clip=paddle.nn.ClipGradByGlobalNorm(1.0))
        opt.enable_flat_arena()
        step = paddle.jit.capture_train_step(lambda *a: loss_fn(net, *a), opt, warmup=3 if mode == "captured" else 10 ** 9)
        for _ in range(6):
            loss = step(*inputs)
        ms = timeit(lambda: step(*inputs), n)
        rows.append((mode, ms, float(loss), step.captured, step.failure))
        if step.failure:
            print(getattr(step, "failure_traceback", ""), flush=True)
    return name, rows


def mlp():
    net = paddle.nn.Sequential(paddle.nn.Linear(784, 512), paddle.nn.ReLU(), paddle.nn.Linear(512, 512), paddle.nn.ReLU(), paddle.nn.Linear(512, 10))
    net.to("gpu")
    ce = paddle.nn.CrossEntropyLoss()
    return net, lambda m, x, y: ce(m(x), y)


def resnet50():
    net = paddle.vision.models.resnet50(num_classes=1000)
    net.to("gpu")
    net = paddle.amp.decorate(net, level="O2", dtype="bfloat16")
    ce = paddle.nn.CrossEntropyLoss()

    def loss_fn(m, x, y):
        with paddle.amp.auto_cast(level="O2", dtype="bfloat16"):
            return ce(m(x).astype("float32"), y)
    return net, loss_fn


def tiny_llama():
    from paddle_b200.models import llama as L

    paddle.set_default_dtype("bfloat16")
    cfg = L.LlamaConfig(vocab_size=8192, hidden_size=1024, intermediate_size=2816, num_hidden_layers=8, num_attention_heads=8, num_key_value_heads=8,
                        max_position_embeddings=512)
    net = L.LlamaForCausalLM(cfg)
    paddle.set_default_dtype("float32")
    return net, lambda m, ids: m(ids[:, :-1], ids[:, 1:])


rng = np.random.RandomState(0)
only = os.environ.get("BENCH_ONLY")
results = [r for r in [
    bench("MLP 784-512-512-10, batch 128, fp32", mlp, (paddle.to_tensor(rng.randn(128, 784).astype("float32")).cuda(), paddle.to_tensor(rng.randint(0, 10, (128,))).cuda()), 200),
    bench("Llama 8 layers h1024 (8 heads x 128) seq 512 batch 4, bf16", tiny_llama, (paddle.to_tensor(rng.randint(0, 8192, (4, 513))).cuda(),), 50),
    bench("ResNet-50 batch 32 224x224, bf16 O2", resnet50, (paddle.to_tensor(rng.randn(32, 3, 224, 224).astype("float32")).cuda().astype("bfloat16"),
                                                            paddle.to_tensor(rng.randint(0, 1000, (32,))).cuda()), 20),
] if r is not None]
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
with open(out_path, "w") as f:
    f.write("# Eager vs captured training step (`paddle.jit.capture_train_step`), one B200, CUDA-event timed, fwd+bwd+clip+AdamW+clear_grad\n\n")
    f.write("| model | mode | ms / step | speed-up | captured | last loss |\n|---|---|---|---|---|---|\n")
    for name, rows in results:
        base = rows[0][1]
        for mode, ms, loss, cap, fail in rows:
            f.write(f"| {name} | {mode} | {ms:.3f} | {base / ms:.2f}x | {cap}{'' if not fail else ' (' + fail[:400].replace(chr(10), ' ') + ')'} | {loss:.4f} |\n")
print(open(out_path).read())
