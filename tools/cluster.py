#!/usr/bin/env python
"""Single-box cluster tool: assign roles to GPUs, template and launch the per-role commands,
collect logs -- the one-box successor of the reference's EC2 launcher.

reference: tools/tf_ec2.py.  Same moving parts, re-targeted from a fleet of EC2 instances
reached over SSH to the 8 GPUs of one box:

===========================  ==========================================================
reference (tf_ec2.py)         here
===========================  ==========================================================
``Cfg`` self-interpolating    :class:`Cfg` -- identical ``%(key)s`` semantics (:17-25)
dict
instance types / spot         ``n_gpus`` (roles are local processes, one GPU each)
requests / NFS mount
``run_tf`` (:445-615)         :func:`run_tf`: setup commands, role assignment (master =
                              task 0, ``worker_<i>`` = task i+1, optional ``ps``,
                              ``evaluator``), substitution of ``PS_HOSTS TASK_ID
                              JOB_NAME WORKER_HOSTS ROLE_ID``, parallel launch with
                              stdout in ``<base_out_dir>/out_<ROLE_ID>``, wall time
                              appended to ``results.txt``, returns ``cluster_save``
``kill_all_python`` /         :func:`kill_all` / :func:`kill_roles`: terminate exactly the
``kill_python``               PIDs recorded in ``<base_out_dir>/pids.json`` (never a pattern)
``list_idle_instances`` /     :func:`list_idle` / :func:`list_running`: GPUs without / with
``list_running_instances``    one of our recorded live processes
``run_command``               :func:`run_command`: run a shell command once per role env
``download_outdir`` /         :func:`download_outdir` / :func:`download_file`: copy from the
``download_file``             run directory (the "NFS") to a local result directory
``launch`` / ``shutdown`` /   :func:`launch` (verify the GPUs exist), :func:`shutdown`
``clean_launch_and_run``      (= kill_all), :func:`clean_launch_and_run`
===========================  ==========================================================
"""
from __future__ import annotations

import json
import os
import shutil
import signal
import subprocess
import sys
import time
from typing import Dict, List, Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from distributedmnist_b200.parallel.launcher import free_port  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from transfer import TransferClient  # noqa: E402  (counterpart of the reference's vendored tools/scp.py)


class Cfg(dict):
    """Dict whose string values (and lists of strings) are ``%``-interpolated against the dict itself."""

    def __getitem__(self, item):
        v = dict.__getitem__(self, item)
        if isinstance(v, list):
            return [x % self if isinstance(x, str) else x for x in v]
        if isinstance(v, str):
            return v % self
        return v


# Default configuration: 1 master + 1 worker + 1 evaluator on one box (reference default: p2.xlarge x4).
cfg = Cfg({
    "name": "Basic",
    "n_masters": 1,                      # should always be 1
    "n_workers": 1,
    "n_ps": 0,                           # accepted; a ps role has nothing to serve and exits
    "n_evaluators": 1,
    "n_gpus": 8,                         # GPUs on the box (0 = CPU/gloo plumbing run)
    "num_replicas_to_aggregate": "2",
    "python": sys.executable,
    "repo": ROOT,
    "base_out_dir": "/tmp/dmnist_runs/%(name)s",
    "setup_commands": ["rm -rf %(base_out_dir)s", "mkdir -p %(base_out_dir)s"],
    "master_pre_commands": [],
    "pre_commands": [],
    "model": "lenet",
    "batch_size": "64",
    "max_steps": "10000",
    "initial_learning_rate": ".1",
    "learning_rate_decay_factor": "0.98",
    "num_epochs_per_decay": "1.0",
    "drop_connect": "False",
    "drop_connect_probability": "0.9",
    "extra_flags": "",
    "eval_extra_flags": "",
    "train_commands": [
        "%(python)s %(repo)s/src/mnist_distributed_train.py "
        "--model=%(model)s --batch_size=%(batch_size)s --max_steps=%(max_steps)s "
        "--initial_learning_rate=%(initial_learning_rate)s "
        "--learning_rate_decay_factor=%(learning_rate_decay_factor)s "
        "--num_epochs_per_decay=%(num_epochs_per_decay)s "
        "--drop_connect=%(drop_connect)s --drop_connect_probability=%(drop_connect_probability)s "
        "--train_dir=%(base_out_dir)s/train_dir --worker_hosts='WORKER_HOSTS' --ps_hosts='PS_HOSTS' "
        "--task_id=TASK_ID --timeline_logging=false "
        "--num_replicas_to_aggregate=%(num_replicas_to_aggregate)s %(extra_flags)s "
        "--job_name=JOB_NAME > %(base_out_dir)s/out_ROLE_ID 2>&1"
    ],
    "evaluate_commands": [
        "sleep 2",
        "%(python)s %(repo)s/src/mnist_eval.py --model=%(model)s --eval_dir=%(base_out_dir)s/eval_dir "
        "--checkpoint_dir=%(base_out_dir)s/train_dir %(eval_extra_flags)s > %(base_out_dir)s/out_evaluator 2>&1",
    ],
})


# ------------------------------------------------------------------------------------------------
def _pidfile(configuration: Cfg) -> str:
    return os.path.join(configuration["base_out_dir"], "pids.json")


def _load_pids(configuration: Cfg) -> Dict[str, Dict]:
    try:
        with open(_pidfile(configuration)) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def _alive(pid: int) -> bool:
    try:
        os.kill(pid, 0)
    except OSError:
        return False
    try:  # a zombie child of ours is not alive
        with open("/proc/%d/stat" % pid) as f:
            return f.read().split(")")[-1].split()[0] != "Z"
    except OSError:
        return False


def n_gpus_available() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:  # noqa: BLE001
        return 0


def launch(argv, configuration: Cfg):
    """Reference ``launch``: make sure the resources of the configuration exist."""
    want = configuration["n_masters"] + configuration["n_workers"]
    have = n_gpus_available()
    if configuration["n_gpus"] and have < min(want, configuration["n_gpus"]):
        print("warning: configuration wants %d GPU replicas, box has %d" % (want, have))
    return have


def run_tf(argv, configuration: Cfg, port: Optional[int] = None) -> Dict:
    """Launch every role of the configuration; returns the ``cluster_save`` dict."""
    assert configuration["n_masters"] == 1
    n_rep = configuration["n_masters"] + configuration["n_workers"]
    port = port or free_port()
    use_gpu = bool(configuration["n_gpus"]) and n_gpus_available() > 0
    worker_host_string = ",".join("127.0.0.1:%d" % (port + 1 + i) for i in range(n_rep))
    ps_host_string = ",".join("127.0.0.1:%d" % (port + 101 + i) for i in range(configuration["n_ps"]))

    for c in configuration["setup_commands"]:
        subprocess.run(c, shell=True, check=False)
    os.makedirs(configuration["base_out_dir"], exist_ok=True)

    def subst(cmd: str, task_id: int, job: str, role: str) -> str:
        return (cmd.replace("PS_HOSTS", ps_host_string).replace("TASK_ID", str(task_id)).replace("JOB_NAME", job)
                .replace("WORKER_HOSTS", worker_host_string).replace("ROLE_ID", role))

    roles: Dict[str, Dict] = {}
    roles["master"] = {"rank": 0, "commands": list(configuration["master_pre_commands"]) + [
        subst(c, 0, "worker", "master") for c in configuration["train_commands"]]}
    for wid in range(configuration["n_workers"]):
        name = "worker_%d" % wid
        roles[name] = {"rank": wid + 1, "commands": list(configuration["pre_commands"]) + [
            subst(c, wid + 1, "worker", name) for c in configuration["train_commands"]]}
    for pid_ in range(configuration["n_ps"]):
        name = "ps_%d" % pid_
        roles[name] = {"rank": None, "commands": list(configuration["pre_commands"]) + [
            subst(c, pid_, "ps", name) for c in configuration["train_commands"]]}
    if configuration["n_evaluators"]:
        assert configuration["n_evaluators"] == 1
        roles["evaluator"] = {"rank": None, "commands": list(configuration["pre_commands"])
                              + list(configuration["evaluate_commands"])}

    start_time = time.time()
    pids: Dict[str, Dict] = {}
    for name, role in roles.items():
        env = dict(os.environ)
        if role["rank"] is not None:
            env.update({"RANK": str(role["rank"]), "LOCAL_RANK": str(role["rank"]), "WORLD_SIZE": str(n_rep),
                        "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
        else:
            for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
                env.pop(k, None)
            if name == "evaluator" and use_gpu:
                # evaluator shares the last GPU when every GPU has a replica, else takes a spare one
                env["CUDA_VISIBLE_DEVICES"] = str(min(n_rep, n_gpus_available() - 1))
        if not use_gpu:
            env["CUDA_VISIBLE_DEVICES"] = ""
        script = " && ".join(role["commands"])
        print("-----------------------\nCommand (%s): %s\n" % (name, script))
        p = subprocess.Popen(["bash", "-c", script], env=env, cwd=configuration["repo"], start_new_session=True)
        pids[name] = {"pid": p.pid, "pgid": p.pid, "gpu": role["rank"] if use_gpu else None}
    with open(_pidfile(configuration), "w") as f:
        json.dump(pids, f)
    run_time = time.time() - start_time
    print("--- %s seconds ---" % run_time)
    with open(os.path.join(configuration["base_out_dir"], "results.txt"), "a") as f:
        f.write("%s\n" % run_time)
    cluster_string = ",".join(sorted(pids))
    return {"configuration": configuration, "name": configuration["name"], "command_machine_assignments": roles,
            "cluster_string": cluster_string, "pids": pids}


def kill_roles(argv, configuration: Cfg, names: Optional[List[str]] = None):
    """Terminate recorded process groups by exact id (reference ``kill_python``)."""
    pids = _load_pids(configuration)
    for name, ent in pids.items():
        if names is not None and name not in names:
            continue
        try:
            os.killpg(ent["pgid"], signal.SIGTERM)
        except OSError:
            pass
    deadline = time.time() + 5
    while time.time() < deadline and any(_alive(e["pid"]) for n, e in pids.items() if names is None or n in names):
        time.sleep(0.1)
    for name, ent in pids.items():
        if (names is None or name in names) and _alive(ent["pid"]):
            try:
                os.killpg(ent["pgid"], signal.SIGKILL)
            except OSError:
                pass
    return pids


def kill_all(argv, configuration: Cfg):
    return kill_roles(argv, configuration, None)


def kill_python(argv, configuration: Cfg):
    if len(argv) != 3:
        print("Usage: python tools/cluster.py kill_python role1,role2,...")
        return None
    return kill_roles(argv, configuration, argv[2].split(","))


def list_running(argv, configuration: Cfg):
    live = {n: e for n, e in _load_pids(configuration).items() if _alive(e["pid"])}
    for n, e in sorted(live.items()):
        print("%s pid=%d gpu=%s" % (n, e["pid"], e["gpu"]))
    return live


def list_idle(argv, configuration: Cfg):
    busy = {e["gpu"] for e in list_running(argv, configuration).values() if e["gpu"] is not None}
    idle = [g for g in range(n_gpus_available()) if g not in busy]
    print("idle GPUs: %s" % idle)
    return idle


def run_command(argv, configuration: Cfg):
    if len(argv) < 4:
        print("Usage: python tools/cluster.py run_command role1,role2 <command>")
        return None
    out = {}
    for name in argv[2].split(","):
        r = subprocess.run(" ".join(argv[3:]), shell=True, capture_output=True, text=True, cwd=configuration["repo"])
        out[name] = r.stdout + r.stderr
        print("%s:\n%s" % (name, out[name]))
    return out


def download_file(argv, configuration: Cfg):
    """``download_file <cluster_string> <file> <outdir>`` -> local path (named ``<cfg name>_<file>``)."""
    _, _, _cluster, fname, outdir = (argv + ["."])[:5]
    os.makedirs(outdir, exist_ok=True)
    src = os.path.join(configuration["base_out_dir"], fname)
    dst = os.path.join(outdir, "%s_%s" % (configuration["name"], os.path.basename(fname)))
    TransferClient().get(src, dst)
    return dst


def download_outdir(argv, configuration: Cfg):
    outdir = argv[3] if len(argv) > 3 else "."
    dst = os.path.join(outdir, configuration["name"])
    if os.path.exists(dst):
        shutil.rmtree(dst)
    TransferClient().get(configuration["base_out_dir"], dst, recursive=True)
    return dst


def shutdown(argv, configuration: Cfg):
    return kill_all(argv, configuration)


def clean_launch_and_run(argv, configuration: Cfg):
    shutdown(argv, configuration)
    launch(argv, configuration)
    return run_tf(argv, configuration)


COMMANDS = {
    "launch": (launch, "Check the box has the GPUs the configuration asks for"),
    "clean_launch_and_run": (clean_launch_and_run, "Kill recorded processes, check resources, launch every role"),
    "shutdown": (shutdown, "Terminate every recorded process"),
    "run_tf": (run_tf, "Launch master/workers/(ps)/evaluator of the configuration"),
    "kill_all_python": (kill_all, "Terminate every recorded process (by exact pid)"),
    "list_idle_instances": (list_idle, "GPUs that run none of our recorded processes"),
    "list_running_instances": (list_running, "Recorded processes still alive"),
    "kill_python": (kill_python, "Terminate the given roles (comma separated)"),
    "run_command": (run_command, "Run a shell command for the given roles"),
    "download_outdir": (download_outdir, "Copy base_out_dir (checkpoints, logs) to a local directory"),
    "download_file": (download_file, "Copy base_out_dir/<file> to a local directory"),
}


def cluster_run(argv, configuration: Cfg):
    if len(argv) < 2 or argv[1] not in COMMANDS:
        print("Usage: python tools/cluster.py [command]\nCommands:")
        for k, (_, h) in COMMANDS.items():
            print("%s - %s" % (k, h))
        return None
    return COMMANDS[argv[1]][0](argv, configuration)


tf_ec2_run = cluster_run   # name used by the reference's benchmark driver

if __name__ == "__main__":
    cluster_run(sys.argv, cfg)
