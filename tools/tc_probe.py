"""Hardware probe of tcgen05 operand layouts (run on the B200 box):  python tools/tc_probe.py [case ...]
Each case runs in its own subprocess (a faulting descriptor must not take the others down)."""
import ctypes
import json
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def smem_desc(lbo, sbo):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def idesc16(M, N, a_fmt, b_fmt):
    """kind::f16 instruction descriptor: formats 0 = f16, 1 = bf16; fp32 accumulate; both operands K-major."""
    return (1 << 4) | (a_fmt << 7) | (b_fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24)


def to16(mat, fmt):
    """fp32 array -> uint16 bit patterns of fp16 (fmt 0) or bf16 (fmt 1); the probe data are small integers (exact)."""
    m = np.ascontiguousarray(mat, dtype=np.float32)
    if fmt == 0:
        return m.astype(np.float16).view(np.uint16)
    return (m.view(np.uint32) >> 16).astype(np.uint16)


def pack_words(u16):
    """uint16 array (even length along the last axis) -> float32 view of the packed 32-bit words."""
    return np.ascontiguousarray(u16).view(np.uint32).view(np.float32)


def kmajor_image16(mat16, lbo, sbo):
    """mat16 [rows][K] uint16 -> canonical no-swizzle K-major image (16-byte chunks = 8 elements)."""
    rows, K = mat16.shape
    size = (rows // 8 - 1) * sbo + (K // 8 - 1) * lbo + 128
    img = np.zeros(size // 2, dtype=np.uint16)
    for r in range(rows):
        for kc in range(K // 8):
            a = ((r % 8) * 16 + (r // 8) * sbo + kc * lbo) // 2
            img[a:a + 8] = mat16[r, 8 * kc:8 * kc + 8]
    return pack_words(img)


def kmajor_image(mat, lbo, sbo):
    """mat [rows][K] fp32 -> canonical no-swizzle K-major image: addr(r,kc) = (r%8)*16 + (r//8)*sbo + kc*lbo."""
    rows, K = mat.shape
    size = (rows // 8 - 1) * sbo + (K // 4 - 1) * lbo + 128
    img = np.zeros(size // 4, dtype=np.float32)
    for r in range(rows):
        for kc in range(K // 4):
            a = ((r % 8) * 16 + (r // 8) * sbo + kc * lbo) // 4
            img[a:a + 4] = mat[r, 4 * kc:4 * kc + 4]
    return img


def run(case):
    import torch
    from sincformer_metacog_speech_enhancement_b200 import _lib
    lib = _lib.load_diag()
    rng = np.random.default_rng(0)
    M = 128
    ri = lambda *s: rng.integers(-4, 5, size=s).astype(np.float32)
    a_from_tmem, a_cols = 0, 0
    if case in ("kmajor_ss", "kmajor_ss_swapped"):
        N, ks = 64, 3
        A, B = ri(M, 8 * ks), ri(N, 8 * ks)
        lbo, sbo = 128, 128 * (8 * ks // 4)
        if case.endswith("swapped"):
            imgA, imgB = kmajor_image(A, lbo, sbo), kmajor_image(B, lbo, sbo)
            dA = dB = smem_desc(sbo, lbo)
        else:
            imgA, imgB = kmajor_image(A, lbo, sbo), kmajor_image(B, lbo, sbo)
            dA = dB = smem_desc(lbo, sbo)
        a_step = b_step = 256
        idc = idesc(M, N, 0, 0)
        expect = A @ B.T
    elif case in ("toeplitz_b", "toeplitz_b_n128", "toeplitz_b_n256"):
        N = {"toeplitz_b": 64, "toeplitz_b_n128": 128, "toeplitz_b_n256": 256}[case]
        ks = 4
        A = ri(M, 8 * ks)
        x = ri(4 * N + 8 * ks + 8)
        Bm = np.stack([x[4 * n:4 * n + 8 * ks] for n in range(N)])       # B[n,k] = x[4n+k]
        lbo, sbo = 128, 128 * (8 * ks // 4)
        imgA, dA, a_step = kmajor_image(A, lbo, sbo), smem_desc(lbo, sbo), 256
        imgB, dB, b_step = x.copy(), smem_desc(16, 128), 32
        idc = idesc(M, N, 0, 0)
        expect = A @ Bm.T
    elif case == "tmem_a_toeplitz_b":
        N, ks = 128, 4
        A = ri(M, 8 * ks)
        x = ri(4 * N + 8 * ks + 8)
        Bm = np.stack([x[4 * n:4 * n + 8 * ks] for n in range(N)])
        imgA, dA, a_step = A.copy(), 0, 8
        a_from_tmem, a_cols = 1, 8 * ks
        imgB, dB, b_step = x.copy(), smem_desc(16, 128), 32
        idc = idesc(M, N, 0, 0)
        expect = A @ Bm.T
    elif case.startswith("h16_"):
        # kind::f16: case = h16_{ts|ss}_{toep|kmaj}_{a fmt}{b fmt}_n{N}   (fmt: h = fp16, b = bf16)
        _, mode, blay, fmts, nn = case.split("_")
        fa, fb = "hb".index(fmts[0]), "hb".index(fmts[1])
        N, ks = int(nn[1:]), 4
        Kt = 16 * ks
        A = ri(M, Kt)
        if blay == "toep":
            x = ri(8 * N + Kt + 16)
            Bm = np.stack([x[8 * n:8 * n + Kt] for n in range(N)])        # B[n,k] = x[8n+k]: row pitch 16 bytes
            imgB, dB, b_step = pack_words(to16(x, fb)), smem_desc(16, 128), 32
        else:
            Bm = ri(N, Kt)
            lbo, sbo = 128, 128 * (Kt // 8)
            imgB, dB, b_step = kmajor_image16(to16(Bm, fb), lbo, sbo), smem_desc(lbo, sbo), 256
        if mode == "ts":
            imgA, dA, a_step = pack_words(to16(A, fa)).reshape(-1).copy(), 0, 8  # [128][Kt/2] packed words, 8 columns per K-step
            a_from_tmem, a_cols = 1, Kt // 2
        else:
            lbo, sbo = 128, 128 * (Kt // 8)
            imgA, dA, a_step = kmajor_image16(to16(A, fa), lbo, sbo), smem_desc(lbo, sbo), 256
        idc = idesc16(M, N, fa, fb)
        expect = A @ Bm.T
    elif case in ("mnmajor_gradfilter", "mnmajor_gradfilter_n256"):
        # D[(c,r), k'] = sum_j g[c, 4j+r] * x[4j+k'];  A MN-major straight from a g tile, B MN-major Toeplitz
        N = 64 if case == "mnmajor_gradfilter" else 256
        ks = 3                                  # 8 j's (32 samples) per step
        P = 32 * ks + 8                         # g tile row pitch in floats (multiple of 4)
        g = ri(32, P)
        x = ri(32 * ks + N + 8)
        Am = np.zeros((M, 8 * ks), np.float32)
        for c in range(32):
            for r in range(4):
                Am[4 * c + r] = g[c, r:r + 32 * ks:4]
        Bm = np.stack([x[kp:kp + 32 * ks:4] for kp in range(N)])         # B[k', j] = x[4j + k']
        imgA, dA, a_step = g.reshape(-1).copy(), smem_desc(128, 4 * P), 128
        imgB, dB, b_step = x.copy(), smem_desc(128, 16), 128
        idc = idesc(M, N, 1, 1)
        expect = Am @ Bm.T
    elif case.startswith("map_"):
        # address-mapping probe: fill the MN-major operand's smem image with word indices (split in two 11-bit
        # digits so every value is exact in TF32), multiply by a K-major one-hot operand, read the mapping off D.
        # case = map_{a|b}_{lbo}_{sbo}[_n{N}]
        parts = case.split("_")
        which, lbo_v, sbo_v = parts[1], int(parts[2]), int(parts[3])
        N, ks, Kt = 64, 1, 8
        words = 4096
        res = {"case": case}
        maps = []
        for digit in (0, 1):
            idx = np.arange(words)
            img = ((idx >> (11 * digit)) & 2047).astype(np.float32)
            eye_rows = M if which == "b" else N
            onehot = np.zeros((eye_rows, Kt), np.float32)
            # logical K index k selected by row k of the K-major operand
            for k in range(Kt):
                onehot[k, k] = 1.0
            imgK = kmajor_image(onehot, 128, 256)
            dK = smem_desc(128, 256)
            dev = torch.device("cuda:0")
            if which == "a":
                tA, tB, dA_, dB_, a_mn, b_mn = img, imgK, smem_desc(lbo_v, sbo_v), dK, 1, 0
            else:
                tA, tB, dA_, dB_, a_mn, b_mn = imgK, img, dK, smem_desc(lbo_v, sbo_v), 0, 1
            tA = torch.from_numpy(np.ascontiguousarray(tA)).to(dev)
            tB = torch.from_numpy(np.ascontiguousarray(tB)).to(dev)
            D = torch.full((M, N), -777.0, device=dev)
            st = torch.zeros(1, dtype=torch.int32, device=dev)
            rc = lib.sinc_tc_probe(ctypes.c_void_p(tA.data_ptr()), tA.numel() * 4, ctypes.c_void_p(tB.data_ptr()),
                                   tB.numel() * 4, dA_, dB_, idesc(M, N, a_mn, b_mn), 1, 0, 0, 0, 0, N,
                                   ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(st.data_ptr()), None)
            torch.cuda.synchronize()
            maps.append(D.cpu().numpy())
            res["rc"], res["status"] = rc, int(st.item())
        if which == "a":
            addr = (maps[0][:, :Kt] + 2048 * maps[1][:, :Kt]).astype(np.int64)      # [m][k] -> word index
        else:
            addr = (maps[0][:Kt, :].T + 2048 * maps[1][:Kt, :].T).astype(np.int64)  # [n][k] -> word index
        res["word_addr_rows0_11"] = addr[:12].tolist()
        res["word_addr_rows_16_32_64"] = [addr[16 if addr.shape[0] > 16 else 0].tolist(), addr[32].tolist(), addr[min(63, addr.shape[0] - 1)].tolist()]
        print(json.dumps(res))
        return res
    elif case.startswith("mn_"):
        # isolate MN-major semantics: one operand MN-major, the other canonical K-major
        N, ks = 64, 3
        Kt = 8 * ks
        lbo, sbo = 128, 128 * (Kt // 4)
        A, Bmat = ri(M, Kt), ri(N, Kt)
        imgA, dA, a_step, a_mn = kmajor_image(A, lbo, sbo), smem_desc(lbo, sbo), 256, 0
        imgB, dB, b_step, b_mn = kmajor_image(Bmat, lbo, sbo), smem_desc(lbo, sbo), 256, 0

        def mn_dense(mat):
            """[rows][Kt] -> MN-major dense image: per k-step block of (rows/4) chunks x 8 k-rows x 16 B."""
            rows = mat.shape[0]
            img = np.zeros(ks * (rows // 4) * 32, np.float32)
            for s_ in range(ks):
                for i in range(rows // 4):
                    for kr in range(8):
                        a = (s_ * (rows // 4) * 128 + i * 128 + kr * 16) // 4
                        img[a:a + 4] = mat[4 * i:4 * i + 4, 8 * s_ + kr]
            return img, (rows // 4) * 128
        if case in ("mn_a_dense", "mn_a_dense_swap"):
            imgA, a_step = mn_dense(A)
            dA = smem_desc(128, 128) if case == "mn_a_dense" else smem_desc(128, 128)
            a_mn = 1
        elif case == "mn_b_dense":
            imgB, b_step = mn_dense(Bmat)
            dB, b_mn = smem_desc(128, 128), 1
        elif case == "mn_a_gtile":
            P = 32 * ks + 8
            g = ri(32, P)
            for c in range(32):
                for r in range(4):
                    A[4 * c + r] = g[c, r:r + 32 * ks:4]
            imgA, dA, a_step, a_mn = g.reshape(-1).copy(), smem_desc(128, 4 * P), 128, 1
        elif case == "mn_b_toeplitz":
            x = ri(32 * ks + N + 8)
            Bmat = np.stack([x[kp:kp + 32 * ks:4] for kp in range(N)])
            imgB, dB, b_step, b_mn = x.copy(), smem_desc(128, 16), 128, 1
        elif case == "mn_b_toeplitz_lbo16":
            x = ri(32 * ks + N + 8)
            Bmat = np.stack([x[kp:kp + 32 * ks:4] for kp in range(N)])
            imgB, dB, b_step, b_mn = x.copy(), smem_desc(16, 16), 128, 1
        else:
            raise SystemExit(f"unknown case {case}")
        idc = idesc(M, N, a_mn, b_mn)
        expect = A @ Bmat.T
    else:
        raise SystemExit(f"unknown case {case}")
    dev = torch.device("cuda:0")
    tA = torch.from_numpy(np.ascontiguousarray(imgA)).to(dev)
    tB = torch.from_numpy(np.ascontiguousarray(imgB)).to(dev)
    D = torch.full((M, N), -777.0, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.sinc_tc_probe(ctypes.c_void_p(tA.data_ptr()), tA.numel() * 4, ctypes.c_void_p(tB.data_ptr()), tB.numel() * 4,
                           dA, dB, idc, ks, a_step, b_step, a_from_tmem, a_cols, N, ctypes.c_void_p(D.data_ptr()),
                           ctypes.c_void_p(st.data_ptr()), None)
    torch.cuda.synchronize()
    got = D.cpu().numpy()
    bad = int((got != expect).sum())
    res = {"case": case, "rc": rc, "status": int(st.item()), "mismatches": bad, "total": int(expect.size),
           "max_abs_err": float(np.abs(got - expect).max())}
    if bad:
        rows = np.unique(np.argwhere(got != expect)[:, 0])
        cols = np.unique(np.argwhere(got != expect)[:, 1])
        res["bad_rows"] = rows[:16].tolist()
        res["bad_cols"] = cols[:16].tolist()
        res["got00"] = got[:2, :4].tolist()
        res["exp00"] = expect[:2, :4].tolist()
    print(json.dumps(res))
    return res


ALL = ["kmajor_ss", "kmajor_ss_swapped", "toeplitz_b", "toeplitz_b_n128", "toeplitz_b_n256", "tmem_a_toeplitz_b",
       "mnmajor_gradfilter", "mnmajor_gradfilter_n256", "mn_a_dense", "mn_b_dense", "mn_a_gtile", "mn_b_toeplitz",
       "mn_b_toeplitz_lbo16"]
H16 = ["h16_ss_kmaj_hh_n64", "h16_ts_kmaj_hh_n64", "h16_ts_toep_hh_n128", "h16_ts_toep_bh_n32", "h16_ts_toep_bb_n112",
       "h16_ts_kmaj_hh_n128", "h16_ss_toep_hh_n256"]

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run(sys.argv[2])
    else:
        for c in (H16 if sys.argv[1:] == ['h16'] else (sys.argv[1:] or ALL)):
            p = subprocess.run([sys.executable, __file__, "--one", c], capture_output=True, text=True, timeout=120)
            out = [l for l in p.stdout.splitlines() if l.startswith("{")]
            print(out[-1] if out else json.dumps({"case": c, "crashed": p.returncode, "stderr": p.stderr[-400:]}), flush=True)
