"""One process of a parameter-server job (role from the TRAINING_ROLE env protocol). Used by test_ps_mode_cpu.py."""
import os
import sys

import numpy as np

import paddle_b200 as paddle
from paddle_b200.distributed import fleet
from paddle_b200.distributed.fleet import ps_mode

out_dir = sys.argv[1]
strategy = fleet.DistributedStrategy()
strategy.a_sync = os.environ.get("PS_ASYNC", "0") == "1"
fleet.init(is_collective=False, strategy=strategy)
assert fleet.server_num() == 1 and fleet.worker_num() == 2

if fleet.is_server():
    fleet.init_server(os.path.join(out_dir, "ckpt") if os.environ.get("PS_LOAD") == "1" else None)
    fleet.run_server()                      # returns when both trainers stopped
    print("server done")
else:
    fleet.init_worker()
    c = ps_mode.client()
    paddle.seed(1)
    emb = ps_mode.DistributedEmbedding("ctr_emb", 8, optimizer="adagrad", lr=0.2)
    c.create_dense_table("bias", (1,), init=[0.0], lr=0.1)
    dense = paddle.nn.Linear(8, 1)
    opt = paddle.optimizer.SGD(0.1, parameters=dense.parameters())
    rng = np.random.RandomState(10 + fleet.worker_index())
    true_w = np.linspace(-1, 1, 50)
    losses = []
    for step in range(60):
        ids = rng.randint(0, 50, size=(16, 3))
        y = paddle.to_tensor((true_w[ids].sum(1, keepdims=True) > 0).astype("float32"))
        bias = paddle.to_tensor(c.pull_dense("bias"))
        bias.stop_gradient = False
        logit = dense(emb(paddle.to_tensor(ids)).sum(1)) + bias
        loss = paddle.nn.functional.binary_cross_entropy_with_logits(logit, y)
        loss.backward()
        c.push_dense("bias", bias.grad.numpy())
        opt.step()
        opt.clear_grad()
        losses.append(float(loss))
    c.flush()
    first, last = np.mean(losses[:10]), np.mean(losses[-10:])
    assert last < first * 0.8, (first, last)
    n_rows = c.table_size("ctr_emb")
    assert 40 <= n_rows <= 50, n_rows
    if fleet.is_first_worker():
        fleet.save_persistables(dirname=os.path.join(out_dir, "ckpt"))
    open(os.path.join(out_dir, f"worker{fleet.worker_index()}.ok"), "w").write(f"{first:.4f} {last:.4f} {n_rows}")
    fleet.stop_worker()
    print("worker done", fleet.worker_index(), first, last)
