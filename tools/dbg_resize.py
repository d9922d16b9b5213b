import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ngp_encode_server_b200 as n
from oracle import oracle as O
w, h, wd, hd = [int(x) for x in sys.argv[1:5]]
fmt = sys.argv[5] if len(sys.argv) > 5 else "rgb24"
nsrc = int(sys.argv[6]) if len(sys.argv) > 6 else 1
text = len(sys.argv) > 7
P = O.Port()
s = n.Session(device=0, max_width=max(w, wd), max_height=max(h, hd), max_sources=nsrc)
m, b = n.synth.load_glyph_table(); s.atlas_set(m, b)
g = O.GlyphTable.load(os.path.join(ROOT, "tests", "golden", "glyphs_aileron20.npz"))
wl = dict(w=w, h=h, wd=wd, hd=hd, fmt=fmt, n_src=nsrc, text="reference" if text else "none")
srcs = n.synth.make_sources(wl, 0)
runs = n.synth.text_runs(wl, 0) if text else None
sc = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"), session=s); dp = n.FrameManager(n.FrameContext(wd, hd, "yuv420p"), session=s)
fin = n.Session.frame_in(fmt, w, h, [(np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(d).reshape(-1), 0, 0) for a, d in srcs])
s.convert(fin, runs, n.api._frame_out(sc, dp))
ws, wdp = O.expected_frame(srcs, fmt, runs, wd, hd, P, g)
def cmp(name, a, b, W, H):
    a = np.frombuffer(a, np.uint8); b = np.frombuffer(b, np.uint8)
    d = np.nonzero(a != b)[0]
    if d.size == 0: print(name, "OK"); return
    ysz = W * H
    print(name, d.size, "bytes differ; first", d[:8], "plane", ["Y" if i < ysz else "UV" for i in d[:4]], "row/col of first", divmod(int(d[0]) if d[0] < ysz else int(d[0]-ysz), W if d[0] < ysz else W//2), "got", a[d[:8]], "want", b[d[:8]])
cmp("scene", sc.cropped(), ws.cropped(), wd, hd)
cmp("depth", dp.cropped(), wdp.cropped(), wd, hd)
print("launches", s.launches, s.last_timing())
