"""A/B of the L2 policy bits of lt3::lbs_fused3_kernel (DPB_LBS_DEBUG: 16 = WITHOUT L2::evict_last on the basis / weight
loads, 32 = WITHOUT st.global.cs vertex stores; the results do not change): bit-compare the vertices and time each variant.
usage: python scripts/lbs_hint_ab.py [B]      (each variant runs in a child with a timeout)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.lbs_ab import CHILD  # noqa: E402


def run(B, dbg, tag):
    out = f'/tmp/lbs_hint_{tag}.pt'
    env = dict(os.environ, DPB_LBS_FUSED='3', DPB_LBS_DEBUG=str(dbg))
    try:
        r = subprocess.run([sys.executable, '-c', CHILD, 'smpl', str(B), out], env=env, timeout=60, capture_output=True, text=True)
    except subprocess.TimeoutExpired:
        return None, 'HANG (killed after 60 s)'
    if r.returncode != 0:
        return None, r.stderr[-600:]
    import torch
    return torch.load(out), ''


if __name__ == '__main__':
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    base = None
    for rnd in range(2):
        for dbg in (0, 16, 32, 48):
            d, err = run(B, dbg, f'{dbg}_{rnd}')
            if d is None:
                print(f'debug={dbg}: {err}', flush=True)
                continue
            if base is None:
                base = d
            diff = float((d['rows'] - base['rows']).abs().max())
            print(f'round {rnd} debug={dbg}: {d["ms"]:.3f} ms  nan={d["nan"]}  max|diff vs debug=0|={diff}', flush=True)
