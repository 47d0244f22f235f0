"""Sensitivity of the attention kernel to its parts (DFD_ATTN_DBG flags of attention_dq.cu; timing only, results are wrong):
1 no exp, 2 no K/V loads after the first fill, 4 no tail MMAs, 8 no P.V, 16 no Q.K^T, 32 no tail TMA boxes, 64 no output store."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch

    sys.path.insert(0, ROOT)
    from dfd import ops

    B, N, H, hd = 64, 729, 16, 72
    qkv = torch.randn(B * N, 3 * H * hd, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        ops.attention_bf16(qkv, B, N, H, hd, impl=5)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.attention_bf16(qkv, B, N, H, hd, impl=5)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"dbg={os.environ.get('DFD_ATTN_DBG', '0'):>4s}  {ts[3]:.3f} ms  {4.0 * N * N * H * hd * B / ts[3] / 1e9:7.1f} TF/s-equivalent")
    if int(os.environ.get("DFD_ATTN_DBG", "0")) & 128:
        import ctypes as C

        from dfd import _lib

        lib = C.CDLL(os.fspath(_lib.LIB_PATH))
        buf = (C.c_longlong * (64 * 16))()
        lib.dfd_debug_read_trace(buf, 64 * 16)
        t0 = buf[0]
        print("  j | softmax: wait_s  ld   math  st   arrive | total || mma: wait_p  pv+commits  qk+commit | gap to next wait")
        for j in range(12):
            r = [buf[j * 16 + k] for k in range(12)]
            print(f"  {j:2d} | start {r[0] - t0:7d}  {r[1] - r[0]:6d} {r[2] - r[1]:5d} {r[3] - r[2]:5d} {r[4] - r[3]:5d} {r[5] - r[4]:5d} | {r[5] - r[0]:6d} || "
                  f"start {r[8] - t0:7d}  {r[9] - r[8]:6d} {r[10] - r[9]:6d} {r[11] - r[10]:6d}")
else:
    for flags in (0, 128, 128 + 127, 128 + 24, 128 + 1):
        env = dict(os.environ, DFD_ATTN_DBG=str(flags))
        subprocess.run([sys.executable, __file__, "child"], env=env)
