import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
import isg_b200
from isg_b200 import _lib, engine
from isg_b200.utils import decode
from oracle import ref_decode as rd
from helpers import unpack_bits
g = np.load('tests/golden/select_points.npz')
dev = torch.device('cuda:0')
for name in 'abcd':
    m = g['in_'+name]; k = int(g['k_'+name]); H, W = m.shape
    md = torch.from_numpy(m).to(dev)
    lib = _lib.lib()
    wsb = int(lib.isg_topk_workspace_bytes(1)); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    thr = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call('isg_topk_threshold', md.data_ptr(), 1, H, W, H*W, k, thr.data_ptr(), ws.data_ptr(), wsb, engine.stream_ptr(dev))
    u = np.array([thr.item()], dtype=np.int32).view(np.uint32)[0]
    f = np.array([u & 0x7FFFFFFF if u & 0x80000000 else ~u], dtype=np.uint32).view(np.float32)[0]
    kth = np.sort(m.ravel())[::-1][k-1]
    hist = ws.view(torch.int32).cpu().numpy().reshape(3, 2048)
    print(name, 'thr gpu', f, 'want', kth, 'hist sums', hist.sum(1))
    out = decode.select_points(md, k).cpu().numpy()
    want = g['out_'+name]
    d = np.argwhere(out != want)
    print('  diffs', len(d), d[:10].tolist())
    for (y, x) in d[:3]:
        y0, y1, x0, x1 = max(0,y-1), min(H,y+2), max(0,x-1), min(W,x+2)
        print('  at', y, x, 'got', out[y,x], 'want', want[y,x]); print(m[y0:y1, x0:x1]); print((m>=kth)[y0:y1, x0:x1].astype(int))
