"""A/B of the fused LBS forward kernels (DPB_LBS_FUSED = 2 | 3): bit-compare the vertices and time both.
usage: python scripts/lbs_ab.py [smpl|smplx] [B]      (each variant runs in a child with a timeout: a hang cannot take the box)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from dposer_b200 import synthetic
from dposer_b200.body_model import BodyModel
mt, B, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
bm = BodyModel(synthetic.make_body_tensors(mt), batch_size=B, model_type=mt).cuda()
inp = {k: v.cuda() for k, v in synthetic.lbs_inputs(B, mt).items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
with torch.no_grad():
    for _ in range(3):
        o = bm(**inp)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = bm(**inp); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    v = o.v
    idx = torch.arange(0, B, max(1, B // 997), device='cuda')
    torch.save({'ms': best, 'sum': float(v.double().sum()), 'abs': float(v.double().abs().sum()), 'rows': v[idx].cpu(),
                'nan': bool(torch.isnan(v).any())}, out)
print('ok', best)
''' % ROOT


def run(mt, B, sel):
    out = f'/tmp/lbs_ab_{sel}.pt'
    env = dict(os.environ, DPB_LBS_FUSED=str(sel))
    try:
        r = subprocess.run([sys.executable, '-c', CHILD, mt, str(B), out], env=env, timeout=40, capture_output=True, text=True)
    except subprocess.TimeoutExpired:
        return None, 'HANG (killed after 40 s)'
    if r.returncode != 0:
        return None, r.stderr[-600:]
    import torch
    return torch.load(out), ''


if __name__ == '__main__':
    mt = sys.argv[1] if len(sys.argv) > 1 else 'smpl'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    res = {}
    for sel in (2, 3):
        d, err = run(mt, B, sel)
        res[sel] = d
        print(f'{mt} B={B} fused={sel}:', f"{d['ms']:.3f} ms  nan={d['nan']} sum={d['sum']:.6f}" if d else err, flush=True)
    if res[2] and res[3]:
        import torch
        diff = (res[2]['rows'] - res[3]['rows']).abs().max()
        print('max |fused3 - fused2| on sampled poses:', float(diff), ' abs-sum equal:', res[2]['abs'] == res[3]['abs'])
