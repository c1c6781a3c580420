#!/usr/bin/env python
"""File transfer between the launcher and the "cluster" -- the one-box counterpart of the reference's vendored SCP
client (reference tools/scp.py: ``SCPClient.put`` :122 / ``SCPClient.get`` :158 over a paramiko transport, used by
tools/tf_ec2.py ``download_file`` / ``download_outdir`` to pull logs and checkpoints off EC2 nodes).

On a single 8-GPU box every role's working directory is a local path, so a transfer is a copy; the client keeps the
reference's call shape (``put(files, remote_path, recursive)``, ``get(remote_path, local_path, recursive)``, a
``progress`` callback receiving ``(name, size, sent)``, ``preserve_times``) so tooling written against it keeps working.
A ``host`` other than this machine is served through the system ``scp`` binary (no paramiko dependency) -- the hook
for multi-box runs.
"""
from __future__ import annotations

import os
import shutil
import socket
import subprocess
from typing import Callable, Iterable, Optional, Union

Progress = Optional[Callable[[str, int, int], None]]


class TransferError(Exception):
    pass


def _is_local(host: Optional[str]) -> bool:
    return host in (None, "", "localhost", "127.0.0.1", socket.gethostname())


class TransferClient:
    def __init__(self, host: Optional[str] = None, user: Optional[str] = None, keyfile: Optional[str] = None,
                 progress: Progress = None, buff_size: int = 1 << 20):
        self.host, self.user, self.keyfile = host, user, keyfile
        self.progress, self.buff_size = progress, buff_size
        self.preserve_times = False

    # ---- local copies -------------------------------------------------------------------------------------------------
    def _copy_file(self, src: str, dst: str) -> None:
        size, sent = os.path.getsize(src), 0
        os.makedirs(os.path.dirname(os.path.abspath(dst)) or ".", exist_ok=True)
        with open(src, "rb") as fi, open(dst, "wb") as fo:
            while True:
                chunk = fi.read(self.buff_size)
                if not chunk:
                    break
                fo.write(chunk)
                sent += len(chunk)
                if self.progress:
                    self.progress(os.path.basename(src), size, sent)
        if self.preserve_times:
            shutil.copystat(src, dst)

    def _copy(self, src: str, dst: str, recursive: bool) -> None:
        if os.path.isdir(src):
            if not recursive:
                raise TransferError("%s is a directory (use recursive=True)" % src)
            target = os.path.join(dst, os.path.basename(os.path.normpath(src))) if os.path.isdir(dst) else dst
            for root, _dirs, files in os.walk(src):
                rel = os.path.relpath(root, src)
                for f in files:
                    self._copy_file(os.path.join(root, f), os.path.normpath(os.path.join(target, rel, f)))
                os.makedirs(os.path.normpath(os.path.join(target, rel)), exist_ok=True)
        elif os.path.exists(src):
            self._copy_file(src, os.path.join(dst, os.path.basename(src)) if os.path.isdir(dst) else dst)
        else:
            raise TransferError("no such file or directory: %s" % src)

    # ---- remote copies (system scp) -------------------------------------------------------------------------------------
    def _remote(self, path: str) -> str:
        return "%s%s:%s" % ((self.user + "@") if self.user else "", self.host, path)

    def _scp(self, src: str, dst: str, recursive: bool) -> None:
        cmd = ["scp", "-q", "-o", "StrictHostKeyChecking=no"] + (["-r"] if recursive else []) + \
              (["-p"] if self.preserve_times else []) + (["-i", self.keyfile] if self.keyfile else []) + [src, dst]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise TransferError("scp failed (%d): %s" % (r.returncode, r.stderr.strip()[-500:]))

    # ---- reference-shaped API ---------------------------------------------------------------------------------------------
    def put(self, files: Union[str, Iterable[str]], remote_path: str = ".", recursive: bool = False,
            preserve_times: bool = False) -> None:
        self.preserve_times = preserve_times
        for f in ([files] if isinstance(files, str) else list(files)):
            if _is_local(self.host):
                self._copy(f, remote_path, recursive)
            else:
                self._scp(f, self._remote(remote_path), recursive)

    def get(self, remote_path: str, local_path: str = "", recursive: bool = False, preserve_times: bool = False) -> None:
        self.preserve_times = preserve_times
        local_path = local_path or os.getcwd()
        if _is_local(self.host):
            self._copy(remote_path, local_path, recursive)
        else:
            self._scp(self._remote(remote_path), local_path, recursive)

    def close(self) -> None:
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
