"""Diagnostic (GPU): mof_spectrum on a synthetic sphere — iterations and time, multigrid cycle against inverse diagonal.
python tests/diag_spectrum.py [level] [count] [vfMode] [cMode]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshopticalflow_b200 import api, synthetic  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    v, t = synthetic.octahedron_sphere(level)
    p = api.default_params()
    p.vfMode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    p.cMode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    for mg in os.environ.get("MGS", "1,0").split(",") if p.vfMode == 0 else ("1",):
        os.environ["MOF_SPECTRUM_MG"] = mg
        al = api.Aligner(0)
        al.set_params(p)
        al.set_mesh(v, t)
        t0 = time.perf_counter()
        try:
            ev, _, its, res = al.spectrum(count, 1e-8, 4000 if mg == "1" else 30000)
            print(f"level {level} V={v.shape[0]} unknowns={al.num_coeffs} vfMode {p.vfMode} preconditioner {('cycle' if p.vfMode == 0 else 'two-cycle' if p.vfMode == 1 else 'diagonal') if mg == '1' else 'diagonal'}: "
                  f"{its} iterations, {time.perf_counter() - t0:.2f} s, residual {res:.2e}, lambda[0] {ev[0]:.8f} lambda[{count - 1}] {ev[-1]:.8f}", flush=True)
        except api.MofError as e:
            print(f"level {level} MG {mg}: {e} after {time.perf_counter() - t0:.2f} s", flush=True)
        al.close()


if __name__ == "__main__":
    main()
