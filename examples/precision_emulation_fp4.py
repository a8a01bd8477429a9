"""Costing study (not a shipped path): what would the MIX precision of the full-rank kernel lose if its two correction products
(hi * lo, lo * hi: today e5m2 `kind::f8f6f4` MMAs, K = 32 per issue slot) ran as FP4 e2m1 block-scaled MMAs (`kind::mxf4`,
K = 64 per slot: 9 -> 7 issue slots per step)?  Re-runs examples/precision_emulation.py with the 8-bit rounding replaced by an
e2m1 rounding, for three scale granularities: 'g' one power-of-two scale per tensor (constant scale factors), 'r' one per row,
'b' one per 32 consecutive K elements (the mx block scale).  Column 'None' is the shipped e5m2 MIX, 'x3' the SPLIT path.
Output: profiles/r2_precision_emulation_fp4.txt.  Verdict in DESIGN.md section 8."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'examples'))
import precision_emulation as pe
GRID = np.array([0,0.5,1,1.5,2,3,4,6],np.float32)
def r4(x, mode):
    """e2m1 rounding of x (any shape, last axis = K).  mode 'g': one power-of-two scale for the whole tensor (max -> <=6);
    'b': one power-of-two scale per 32 consecutive K elements (mx block scaling); 'r': one per row (all K)"""
    x = x.astype(np.float32)
    a = np.abs(x)
    if mode == 'g':
        m = a.max()
        s = np.float32(2.0**np.ceil(np.log2(max(m,1e-30)/6.0)))
        s = np.broadcast_to(s, x.shape)
    else:
        K = x.shape[-1]
        blk = 32 if mode == 'b' else K
        xs = a.reshape(x.shape[:-1]+(K//blk, blk))
        m = xs.max(-1, keepdims=True)
        s = 2.0**np.ceil(np.log2(np.maximum(m,1e-30)/6.0))
        s = np.broadcast_to(s, xs.shape).reshape(x.shape).astype(np.float32)
    y = a / s
    idx = np.abs(y[...,None]-GRID).argmin(-1)
    return (np.sign(x)*GRID[idx]*s).astype(np.float32)
orig_r8 = pe.r8
MODE = {'m': None}
def r8(x, fmt, s=1.0):
    if MODE['m'] is None: return orig_r8(x, fmt, s)
    return r4(x, MODE['m'])
pe.r8 = r8
if __name__ == '__main__':
    cases = [('reference', None, 1, 64), ('trained', 0.3, 1, 64), ('trained', 0.3, 4, 64), ('trained', 1.0, 1, 64), ('trained',0.3,1,16), ('trained',0.3,1,128)]
    for style, es, ws, H in cases:
        row = {}
        for m in (None, 'g', 'r', 'b'):
            MODE['m'] = m
            o = pe.run(style, emb_std=es, wscale=ws, B=128, H=H)
            row[str(m)] = o['mix55']
            row['x3'] = o['x3']; row['rho'] = o['rho']
        print(style, es, ws, H, {k: f"{v:.2e}" for k, v in row.items()}, flush=True)
