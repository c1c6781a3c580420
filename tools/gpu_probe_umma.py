"""Hardware probe (see csrc/probes/umma_probe.cu): UMMA descriptors starting at a row offset inside a swizzled tile."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from distributedmnist_b200.ops.lib import check, load, ptr, stream_ptr  # noqa: E402


def main():
    import ctypes
    from distributedmnist_b200.ops.build import build_probes
    lib = ctypes.CDLL(build_probes())       # the probes live in their own library (csrc/probes/)
    os.makedirs("gpurun_out", exist_ok=True)
    lines = []
    torch.manual_seed(0)
    for row_bytes in (128, 64):
        K = row_bytes // 2
        A = (torch.randn(144, K, device="cuda")).to(torch.bfloat16)
        Bm = (torch.randn(64, K, device="cuda")).to(torch.bfloat16)
        for mode in (0, 1):
            errs = []
            for shift in range(0, 17):
                out = torch.zeros(128, 64, device="cuda")
                check(lib.dm_umma_shift_probe(ptr(A), ptr(Bm), ptr(out), row_bytes, shift, mode, stream_ptr()), "probe")
                torch.cuda.synchronize()
                ref = A[shift:shift + 128].float() @ Bm.float().t()
                errs.append((out - ref).abs().max().item())
            lines.append("row_bytes=%d mode=%d (base_offset=%s): max err per row shift 0..16: %s" % (
                row_bytes, mode, "0" if mode == 0 else "(addr>>7)&7", " ".join("%.2g" % e for e in errs)))
    txt = "\n".join(lines)
    print(txt)
    open("gpurun_out/umma_probe.txt", "w").write(txt + "\n")


if __name__ == "__main__":
    main()
