"""GPU timing of isg_box_nms alone (CUDA events, back-to-back launches on one stream): the parallel suppression scan
(ISG_NMS_ROUNDS=32, default) against the sequential scan (ISG_NMS_ROUNDS=0), fused small kernel and staged path.
  python tools/nms_timing.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isg_b200  # noqa
from isg_b200 import _lib, engine, synth


def case(B, n, cap, extent, thr, ncls, iters=300):
    dev = torch.device("cuda", 0)
    dets = [synth.make_nms_boxes(100 + b, n, extent=extent, thr=thr, plus1=False) for b in range(B)]
    boxes = torch.zeros((B, cap, 4), dtype=torch.float32)
    scores = torch.zeros((B, cap), dtype=torch.float32)
    cls = torch.zeros((B, cap), dtype=torch.int32)
    cnt = torch.zeros(B, dtype=torch.int32)
    rs = np.random.RandomState(7)
    for b, d in enumerate(dets):
        m = min(len(d), cap)
        boxes[b, :m] = torch.from_numpy(d[:m, :4]); scores[b, :m] = torch.from_numpy(d[:m, 4])
        cls[b, :m] = torch.from_numpy(rs.randint(0, ncls, size=m).astype(np.int32)); cnt[b] = m
    boxes, scores, cls, cnt = boxes.to(dev), scores.to(dev), cls.to(dev), cnt.to(dev)
    keep = torch.empty((B, cap), dtype=torch.int32, device=dev); nk = torch.empty(B, dtype=torch.int32, device=dev)
    wsb = int(_lib.lib().isg_box_nms_workspace_bytes(B, cap))
    ws, ws_ptr = engine.aligned_workspace(max(wsb, 256), dev)
    st = engine.stream_ptr(dev)
    out = {}
    for rounds in ("0", "32"):
        os.environ["ISG_NMS_ROUNDS"] = rounds
        _lib.lib().isg_debug_reload_tuning()
        run = lambda: _lib.call("isg_box_nms", boxes.data_ptr(), scores.data_ptr(), cls.data_ptr(), 0, cnt.data_ptr(), B, cap, thr,
                                _lib.ISG_NMS_TV_GT, keep.data_ptr(), nk.data_ptr(), ws_ptr, wsb, st)
        for _ in range(20): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): run()
        e1.record(); torch.cuda.synchronize()
        out[rounds] = (e0.elapsed_time(e1) / iters * 1e3, keep.clone(), nk.clone())
    os.environ.pop("ISG_NMS_ROUNDS"); _lib.lib().isg_debug_reload_tuning()
    same = bool(torch.equal(out["0"][2], out["32"][2])) and all(
        torch.equal(out["0"][1][b, :int(out["0"][2][b])], out["32"][1][b, :int(out["32"][2][b])]) for b in range(B))
    print("B=%d n=%d cap=%d classes=%d: sequential scan %.1f us, parallel rounds %.1f us per call (kept %s, identical: %s)" % (
        B, int(cnt[0]), cap, ncls, out["0"][0], out["32"][0], out["32"][2].cpu().numpy().tolist()[:4], same), flush=True)


if __name__ == "__main__":
    case(8, 256, 1024, 1400.0, 0.5, 8)        # the bench step's box branch (fused small kernel)
    case(8, 1000, 1024, 2500.0, 0.5, 8)       # small kernel at its capacity
    case(4, 850, 2048, 2500.0, 0.5, 8)        # crowd config: staged path, whole matrix in shared memory
    case(1, 1000, 1000, 2500.0, 0.5, 80)      # config-5 sized candidate set
    case(1, 3000, 3000, 4000.0, 0.5, 8)       # beyond the shared-memory matrix: sequential chunks
