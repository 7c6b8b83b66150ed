#!/usr/bin/env python3
"""Quick device-time probe of the large-component path (config C4 shape): python tools/probe_large.py [ndof] [E] [L]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pasif_b200 import capi, frontend, workloads as wl


class mod:
    coo_tensor = frontend.coo_tensor


def flops(s, j):
    def nnzk(d):
        return len(set(zip(list(map(int, d["B"][1])) + list(map(int, d["K"][1])), list(map(int, d["B"][2])) + list(map(int, d["K"][2])))))
    def D(d, force):
        return 2 * nnzk(d) + 3 * len(d["Gamma"][1]) + 4 * len(d["Lambda"][1]) + 5 * len(d.get("Psi", ([], [], []))[1]) + (2 * int(np.count_nonzero(d["F"])) if force else 0)
    return 4 * D(s, True) + 13 * s["n"], 4 * D(j, False) + 13 * j["n"] + 2 * s["n"]


def main():
    ndof = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    exc = wl.random_sine_excitations(E, L, 16000, seed=5)
    for dense in (False, True):
        s = wl.chain_system(ndof, dense=dense)
        j = wl.chain_jacobian(s, 4)
        ff, fb = flops(s, j)
        for prec in ("f64", "f32"):
            for opts in ([{}] if not dense else ([{}, {"gemm": "tile"}] if prec == "f64" else [{}, {"tcgen05": "off"}, {"gemm": "tile"}])):
                p = frontend.make_frontend(capi.engine_class(prec))(exc, 16000, 0)
                for k, v in opts.items():
                    p._setOption(k, v)
                t0 = time.time()
                p.setSystems(*wl.to_frontend_args([s], mod))
                p.setJacobian(*wl.to_frontend_args([j], mod, with_psi=True))
                ts = time.time() - t0
                p._setOption("profile", "on")
                res = {}
                for rep in range(2):
                    p._getAmplitudes(fetch=False)
                    res["fwd_ms"] = p._query("last_device_ms")
                    p._getGradient(0, fetch=False)
                    res["grad_ms"] = p._query("last_device_ms")
                    res["gfwd_ms"], res["bwd_ms"] = p._query("ms_gfwd"), p._query("ms_bwd")
                p._setOption("profile", "off")
                for rep in range(2):
                    p._getAmplitudes(fetch=False)
                    res["fwd_ovl_ms"] = p._query("last_device_ms")
                    p._getGradient(0, fetch=False)
                    res["grad_ovl_ms"] = p._query("last_device_ms")
                nsp = int(np.sqrt(L)); c = int(np.ceil(L / nsp))
                fwd_tf = E * L * ff / (res["fwd_ms"] * 1e-3) / 1e12
                gf_tf = E * (2 * L - c) * ff / (res["gfwd_ms"] * 1e-3) / 1e12
                bw_tf = E * (L - c) * fb / (res["bwd_ms"] * 1e-3) / 1e12
                sz = 4 if prec == "f32" else 8
                NR = (s["n"] + 31) // 32 * 32
                bw_gbs = E * (L - c) * NR * sz / (res["bwd_ms"] * 1e-3) / 1e9
                print(f"ndof={ndof} dense={dense} {prec} {opts} N={s['n']} NA={j['n']} E={E} L={L} setup={ts:.1f}s gemm_classes={int(p._query('large_gemm_classes'))} "
                      f"F_fwd={ff} F_bwd={fb} | fwd {res['fwd_ms']:.1f} ms = {E*L/(res['fwd_ms']*1e-3):.3e} traj-steps/s {fwd_tf:.2f} TF | "
                      f"OVERLAPPED fwd {res['fwd_ovl_ms']:.1f} ms = {E*L/(res['fwd_ovl_ms']*1e-3):.3e} traj-steps/s, grad {res['grad_ovl_ms']:.1f} ms = {E/(res['grad_ovl_ms']*1e-3):.3e} evals/s | "
                      f"grad {res['grad_ms']:.1f} ms gfwd {res['gfwd_ms']:.1f} ({gf_tf:.2f} TF) bwd {res['bwd_ms']:.1f} ({bw_tf:.2f} TF, {bw_gbs:.0f} GB/s read)", flush=True)
                p.close() if hasattr(p, "close") else None


if __name__ == "__main__":
    main()
