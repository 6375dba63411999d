from __future__ import annotations

import argparse
import os
import signal
import socket
import subprocess
import sys
import time


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def parse(argv=None):
    ap = argparse.ArgumentParser("paddle_b200.distributed.launch")
    ap.add_argument("--master", default=None, help="ip:port of the rendezvous endpoint (default 127.0.0.1:<free port>)")
    ap.add_argument("--nnodes", default="1")
    ap.add_argument("--rank", type=int, default=0, help="node rank")
    ap.add_argument("--nproc_per_node", type=int, default=None)
    ap.add_argument("--devices", "--gpus", dest="devices", default=None, help="comma separated device ids")
    ap.add_argument("--log_dir", default="log")
    ap.add_argument("--job_id", default="default")
    ap.add_argument("--run_mode", default="collective")
    ap.add_argument("--max_restart", type=int, default=0)
    ap.add_argument("--elastic_timeout", type=int, default=30)
    ap.add_argument("--server_num", type=int, default=None, help="parameter-server mode: servers started on this node")
    ap.add_argument("--trainer_num", type=int, default=None, help="parameter-server mode: trainers started on this node")
    ap.add_argument("--servers", default="", help="parameter-server mode: explicit ip:port list of the servers")
    ap.add_argument("--auto_tuner_json", default=None, help="auto-tuner mode: JSON with the model / search space (distributed/auto_tuner.py); every trial is one "
                    "launch of the training script with the candidate in B200_TUNE_* variables; the script prints a JSON line with the metric")
    ap.add_argument("training_script")
    ap.add_argument("training_script_args", nargs=argparse.REMAINDER)
    return ap.parse_args(argv)


def launch_ps(args):
    """Parameter-server controller: starts the servers and trainers of this node with the TRAINING_ROLE / PADDLE_PSERVERS_IP_PORT_LIST
    protocol (Parity: launch/controllers/ps.py). A failing process tears the job down; servers exit by themselves when every trainer
    called fleet.stop_worker()."""
    n_srv = args.server_num or (len(args.servers.split(",")) if args.servers else 1)
    n_trn = args.trainer_num or 1
    endpoints = args.servers.split(",") if args.servers else [f"127.0.0.1:{_free_port()}" for _ in range(n_srv)]
    os.makedirs(args.log_dir, exist_ok=True)
    base = dict(os.environ, PADDLE_PSERVERS_IP_PORT_LIST=",".join(endpoints), PADDLE_TRAINERS_NUM=str(n_trn), PADDLE_JOB_ID=args.job_id)
    procs = []
    for i in range(n_srv):
        ip, port = endpoints[i].rsplit(":", 1)
        env = dict(base, TRAINING_ROLE="PSERVER", POD_IP=ip, PADDLE_PORT=port, PADDLE_PSERVER_ID=str(i))
        log = open(os.path.join(args.log_dir, f"serverlog.{i}"), "w")
        procs.append((subprocess.Popen([sys.executable, "-u", args.training_script, *args.training_script_args], env=env, stdout=log, stderr=subprocess.STDOUT), log))
    for i in range(n_trn):
        env = dict(base, TRAINING_ROLE="TRAINER", PADDLE_TRAINER_ID=str(i))
        log = open(os.path.join(args.log_dir, f"workerlog.{i}"), "w")
        procs.append((subprocess.Popen([sys.executable, "-u", args.training_script, *args.training_script_args], env=env,
                                       stdout=log if i != 0 else None, stderr=subprocess.STDOUT if i != 0 else None), log))
    failed = None
    while True:
        alive = False
        for p, _ in procs:
            rc = p.poll()
            if rc is None:
                alive = True
            elif rc != 0 and failed is None:
                failed = rc
        if failed is not None or not alive:
            break
        time.sleep(0.3)
    if failed is not None:
        for p, _ in procs:
            if p.poll() is None:
                p.send_signal(signal.SIGTERM)
        for p, _ in procs:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
        print(f"[launch] ps job failed with exit code {failed}; see {args.log_dir}/serverlog.* and workerlog.*", file=sys.stderr)
    for _, log in procs:
        log.close()
    return failed or 0


def launch(args):
    if args.run_mode == "ps" or args.server_num or args.servers:
        return launch_ps(args)
    import torch

    devices = [d for d in args.devices.split(",")] if args.devices else [str(i) for i in range(max(1, torch.cuda.device_count()))]
    nproc = args.nproc_per_node or len(devices)
    nnodes = int(str(args.nnodes).split(":")[0])
    if args.master:
        addr, port = args.master.split(":")
    else:
        addr, port = "127.0.0.1", str(_free_port())
    os.makedirs(args.log_dir, exist_ok=True)
    world = nproc * nnodes
    restarts = 0
    while True:
        procs = []
        for lr in range(nproc):
            rank = args.rank * nproc + lr
            env = dict(os.environ)
            env.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(lr), "MASTER_ADDR": addr, "MASTER_PORT": port,
                        "PADDLE_TRAINER_ID": str(rank), "PADDLE_TRAINERS_NUM": str(world), "PADDLE_RANK_IN_NODE": str(lr),
                        "PADDLE_LOCAL_SIZE": str(nproc), "PADDLE_JOB_ID": args.job_id, "FLAGS_selected_gpus": devices[lr % len(devices)]})
            log = open(os.path.join(args.log_dir, f"workerlog.{rank}"), "w")
            p = subprocess.Popen([sys.executable, "-u", args.training_script, *args.training_script_args], env=env,
                                 stdout=log if lr != 0 else None, stderr=subprocess.STDOUT if lr != 0 else None)
            procs.append((p, log))
        failed = None
        while True:   # watcher: first failure tears the pod down
            alive = False
            for p, _ in procs:
                rc = p.poll()
                if rc is None:
                    alive = True
                elif rc != 0 and failed is None:
                    failed = rc
            if failed is not None or not alive:
                break
            time.sleep(0.5)
        if failed is not None:
            for p, _ in procs:
                if p.poll() is None:
                    p.send_signal(signal.SIGTERM)
            for p, _ in procs:
                try:
                    p.wait(timeout=10)
                except subprocess.TimeoutExpired:
                    p.kill()
        for _, log in procs:
            log.close()
        if failed is None:
            return 0
        restarts += 1
        if restarts > args.max_restart:
            print(f"[launch] worker failed with exit code {failed}; see {args.log_dir}/workerlog.*", file=sys.stderr)
            return failed
        print(f"[launch] restart {restarts}/{args.max_restart} after failure (exit code {failed})", file=sys.stderr)
        port = str(_free_port()) if not args.master else port


def launch_auto_tuner(args):
    """`--auto_tuner_json cfg.json`: trial loop of distributed.auto_tuner.  Parity: the auto-tuner mode of paddle.distributed.launch
    (python/paddle/distributed/launch/controllers/collective.py + auto_tuner/tuner.py)."""
    import json

    from ..auto_tuner import AutoTuner

    with open(args.auto_tuner_json) as f:
        cfg = json.load(f)
    nproc = args.nproc_per_node or cfg.get("num_gpus", 1)
    cfg.setdefault("num_gpus", nproc * int(str(args.nnodes).split(":")[0]))
    os.makedirs(args.log_dir, exist_ok=True)
    cfg.setdefault("history_path", os.path.join(args.log_dir, "auto_tuner_history.jsonl"))
    tuner = AutoTuner(cfg)
    cmd = [sys.executable, "-m", "paddle_b200.distributed.launch", "--nproc_per_node", str(nproc), "--log_dir", os.path.join(args.log_dir, "trial")]
    if args.master:
        cmd += ["--master", args.master]
    cmd += [args.training_script] + list(args.training_script_args)
    best = tuner.tune(command=cmd, max_trials=int(cfg.get("max_trials", 8)), max_time_s=cfg.get("max_time_s"), timeout_s=int(cfg.get("trial_timeout_s", 1800)),
                      metric_key=cfg.get("metric_key", "value"))
    tuner.recorder.to_csv(os.path.join(args.log_dir, "auto_tuner_history.csv"))
    print(json.dumps({"auto_tuner_best": best, "trials": len(tuner.history)}))
    return 0 if best is not None else 1


def main(argv=None):
    args = parse(argv)
    sys.exit(launch_auto_tuner(args) if args.auto_tuner_json else launch(args))
