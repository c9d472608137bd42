"""Exercise the compiled binding's error paths one per subprocess (a crash must not hide the others)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["ok", "template", "dtype", "shape", "rc"]


def one(case):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import faulthandler
    faulthandler.enable()
    import torch
    import flute_b200 as flute
    from flute_b200 import ops, utils
    from helpers import make_case
    dev = torch.device("cuda", 0)
    ws = utils.get_workspace_streamk(dev)
    c = make_case(6, 1024, 512, 4, 64, "bfloat16", seed=21)
    Q, S, table, t2 = [c[k].to(dev) for k in ("Q", "S", "table", "table2")]
    x = c["A"].to(dev)
    args = dict(ok=(x, Q, S, table, t2, ws, 4, 64, 0, 148), template=(x, Q, S, table, t2, ws, 4, 64, 9999, 148),
                dtype=(x.float(), Q, S, table, t2, ws, 4, 64, 0, 148), shape=(x, Q[:-1], S, table, t2, ws, 4, 64, 0, 148),
                rc=(x, Q, S, table, t2, ws[:1024], 4, 64, 0, 148))[case]
    try:
        out = flute.qgemm(*args)
        torch.cuda.synchronize()
        print(f"[{ops.BINDING}] {case}: returned {tuple(out.shape)}", flush=True)
    except Exception as e:
        print(f"[{ops.BINDING}] {case}: raised {type(e).__name__}: {str(e)[:90]}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1])
    else:
        for env in ({}, {"FLUTE_B200_PY_OPS": "1"}):
            for case in CASES:
                r = subprocess.run([sys.executable, __file__, case], env=dict(os.environ, **env), capture_output=True, text=True, timeout=120)
                tail = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
                print(f"rc={r.returncode} {tail}" + ("" if r.returncode == 0 else " | " + " / ".join(r.stderr.strip().splitlines()[:3])), flush=True)
