"""Bring-up probe for the 8-bit / 4-bit tensor-core kinds (fx_dbg_bs_tile, libflux_b200_dbg.so): one 128 x N tile per
launch against a float64 reference.  Prints the relative error of every variant (and of scale-free controls) so that one
GPU run tells which operand / scale-factor layouts the hardware expects.  Test infrastructure only."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flux-generator_b200"))
from flux import _native  # noqa: E402

dev = "cuda"
lib = _native.dbg_lib()
E2M1 = torch.tensor([0, .5, 1, 1.5, 2, 3, 4, 6, -0., -.5, -1, -1.5, -2, -3, -4, -6], dtype=torch.float64)


def e4m3_bytes(shape, g, scale=1.0):
    x = (torch.randn(shape, generator=g) * scale).to(torch.float8_e4m3fn)
    return x.view(torch.uint8), x.to(torch.float64)


def fp4_bytes(rows, kbytes, g):
    b = torch.randint(0, 256, (rows, kbytes), generator=g, dtype=torch.uint8)
    lo, hi = (b & 15).long(), (b >> 4).long()
    vals = torch.stack([E2M1[lo], E2M1[hi]], dim=-1).reshape(rows, 2 * kbytes)  # element 2i = low nibble
    return b, vals


def ue8m0(shape, g, unit=False):
    b = torch.full(shape, 127, dtype=torch.uint8) if unit else torch.randint(121, 132, shape, generator=g, dtype=torch.uint8)
    return b, torch.pow(2.0, b.double() - 127)


def ue4m3(shape, g, unit=False):
    if unit:
        x = torch.ones(shape).to(torch.float8_e4m3fn)
    else:
        x = (torch.rand(shape, generator=g) * 3.5 + 0.25).to(torch.float8_e4m3fn)
    return x.view(torch.uint8), x.to(torch.float64)


def run(A, B, SFA, SFB, N, kbytes, kind, a_tmem=0, b_mn=0, nsf=0, b_lbo=0, b_sbo=0, b_kstep=0, cp_lbo=0, cp_sbo=128):
    D = torch.full((128, N), float("nan"), device=dev, dtype=torch.float32)
    a, b = A.contiguous().to(dev), B.contiguous().to(dev)
    sfa = SFA.contiguous().to(dev) if SFA is not None else None
    sfb = SFB.contiguous().to(dev) if SFB is not None else None
    rc = lib.fx_dbg_bs_tile(a.data_ptr(), b.data_ptr(), None if sfa is None else sfa.data_ptr(), None if sfb is None else sfb.data_ptr(),
                            D.data_ptr(), N, kbytes, kind, a_tmem, b_mn, nsf, b_lbo, b_sbo, b_kstep, cp_lbo, cp_sbo, _native.stream())
    _native.check(rc)
    torch.cuda.synchronize()
    return D.double().cpu()


def rel(d, ref):
    return ((d - ref).norm() / ref.norm()).item() if torch.isfinite(d).all() else float("nan")


def main():
    g = torch.Generator().manual_seed(1)
    # ---- kind 0: plain FP8
    for N, kb in ((128, 128), (256, 256)):
        Ab, Af = e4m3_bytes((128, kb), g)
        Bb, Bf = e4m3_bytes((N, kb), g)
        ref = Af @ Bf.T
        print(f"kind0 f8 SS            N={N} K={kb}: rel {rel(run(Ab, Bb, None, None, N, kb, 0), ref):.2e}", flush=True)
        print(f"kind0 f8 A-from-TMEM   N={N} K={kb}: rel {rel(run(Ab, Bb, None, None, N, kb, 0, a_tmem=1), ref):.2e}", flush=True)
    Ab, Af = e4m3_bytes((128, 128), g)
    Bb, Bf = e4m3_bytes((128, 128), g)           # B [K=128][N=128]
    ref = Af @ Bf
    for lbo, sbo, ks in ((16384, 1024, 4096), (0, 1024, 4096), (1024, 1024, 4096), (16, 1024, 4096), (16384, 1024, 2048), (1024, 128, 4096)):
        for at in (0, 1):
            print(f"kind0 f8 B MN-major    lbo={lbo} sbo={sbo} kstep={ks} a_tmem={at}: rel "
                  f"{rel(run(Ab, Bb, None, None, 128, 128, 0, a_tmem=at, b_mn=1, b_lbo=lbo, b_sbo=sbo, b_kstep=ks), ref):.2e}", flush=True)
    # ---- kind 1: MXFP8 (UE8M0 per 32)
    for N, kb in ((128, 128), (256, 256)):
        Ab, Af = e4m3_bytes((128, kb), g)
        Bb, Bf = e4m3_bytes((N, kb), g)
        nsf = kb // 32
        for unit in (True, False):
            sa_b, sa_f = ue8m0((128, nsf), g, unit)
            sb_b, sb_f = ue8m0((N, nsf), g, unit)
            ref = (Af * sa_f.repeat_interleave(32, 1)) @ (Bf * sb_f.repeat_interleave(32, 1)).T
            for cl, cs in ((0, 128), (16, 128), (128, 128), (128, 0)):
                print(f"kind1 mxf8 unit_sf={int(unit)} N={N} K={kb} cp(lbo={cl},sbo={cs}): rel "
                      f"{rel(run(Ab, Bb, sa_b, sb_b, N, kb, 1, nsf=nsf, cp_lbo=cl, cp_sbo=cs), ref):.2e}", flush=True)
    # ---- kind 2: NVFP4 (e2m1, UE4M3 per 16), kind 3: MXFP4 (UE8M0 per 32)
    for kind, vec, sf in ((2, 16, ue4m3), (3, 32, ue8m0)):
        for N, kb in ((128, 128), (256, 256)):
            Ab, Af = fp4_bytes(128, kb, g)
            Bb, Bf = fp4_bytes(N, kb, g)
            nsf = 2 * kb // vec
            for unit in (True, False):
                sa_b, sa_f = sf((128, nsf), g, unit)
                sb_b, sb_f = sf((N, nsf), g, unit)
                ref = (Af * sa_f.repeat_interleave(vec, 1)) @ (Bf * sb_f.repeat_interleave(vec, 1)).T
                print(f"kind{kind} fp4 vec{vec} unit_sf={int(unit)} N={N} K={2 * kb}: rel "
                      f"{rel(run(Ab, Bb, sa_b, sb_b, N, kb, kind, nsf=nsf), ref):.2e}", flush=True)


def main_pair():
    """CTA pair (cta_group::2) NVFP4: M = 256, which shared memory / TMEM do the scale factors of a pair come from?"""
    g = torch.Generator().manual_seed(2)
    for N, kb in ((256, 128), (192, 256), (128, 128)):
        Ab, Af = fp4_bytes(256, kb, g)
        Bb, Bf = fp4_bytes(N, kb, g)
        nsf = 2 * kb // 16
        for unit in (True, False):
            sa_b, sa_f = ue4m3((256, nsf), g, unit)
            sb_b, sb_f = ue4m3((N, nsf), g, unit)
            ref = (Af * sa_f.repeat_interleave(16, 1)) @ (Bf * sb_f.repeat_interleave(16, 1)).T
            for mode in (0, 1):
                D = torch.full((256, N), float("nan"), device=dev, dtype=torch.float32)
                t_ = [x.contiguous().to(dev) for x in (Ab, Bb, sa_b, sb_b)]
                _native.check(lib.fx_dbg_bs2_tile(t_[0].data_ptr(), t_[1].data_ptr(), t_[2].data_ptr(), t_[3].data_ptr(), D.data_ptr(),
                                                  N, kb, nsf, mode, _native.stream()))
                torch.cuda.synchronize()
                d = D.double().cpu()
                print(f"pair nvfp4 unit_sf={int(unit)} N={N} K={2 * kb} sfb_mode={mode}: rel all {rel(d, ref):.2e}  "
                      f"rows 0-127 {rel(d[:128], ref[:128]):.2e}  rows 128-255 {rel(d[128:], ref[128:]):.2e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pair":
        main_pair()
    else:
        main()
