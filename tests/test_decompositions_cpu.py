"""The algebra behind three CUDA paths, checked in float64 on the CPU with plain torch ops -- the index conventions here are the
ones the kernels implement (csrc/conv_k2r.cu, csrc/conv_stem.cu, sp_axis() in csrc/conv_tc.cu):

* kernel-to-row RGB tail: conv3x3(cat(up2x(u), v)) = gather of a 1x1 problem at SOURCE resolution + direct 3x3 over v, and its
  backward through the D rows;
* space-to-depth stem: conv7x7 / stride 2 / pad 3 = conv4x4 / stride 1 over the 2x2 space-to-depth image with 2 cells of padding in
  front (and the weight-gradient gather back into [co][7][7][c]);
* sub-pixel decomposition: a k x k convolution over a nearest-2x-upsampled source = per output parity class a smaller convolution
  over the source whose taps are SUMS of the original taps.
"""
import torch
import torch.nn.functional as F

DT = torch.float64


def test_kernel_to_row_tail_identities():
    torch.manual_seed(0)
    n, cu, cs, co, hs, ws = 2, 8, 3, 3, 5, 6
    h, w = 2 * hs, 2 * ws
    u = torch.randn(n, cu, hs, ws, dtype=DT)
    v = torch.randn(n, cs, h, w, dtype=DT)
    W = torch.randn(co, cu + cs, 3, 3, dtype=DT)
    u_ = u.clone().requires_grad_(True)
    W_ = W.clone().requires_grad_(True)
    y = F.conv2d(torch.cat([F.interpolate(u_, scale_factor=2, mode="nearest"), v], 1), W_, padding=1)
    # forward: Z[s][tap, co] once per source pixel; child (a, b) of s takes tap (tr, tc) from neighbour (floor((a+tr-1)/2), floor((b+tc-1)/2))
    Z = torch.einsum("nchw,octs->nhwtso", u, W[:, :cu])
    yy = F.conv2d(v, W[:, cu:], padding=1)
    for sy in range(hs):
        for sx in range(ws):
            for a in range(2):
                for b in range(2):
                    for tr in range(3):
                        for tc in range(3):
                            ny, nx = sy + (a + tr + 1) // 2 - 1, sx + (b + tc + 1) // 2 - 1
                            if 0 <= ny < hs and 0 <= nx < ws:
                                yy[:, :, 2 * sy + a, 2 * sx + b] += Z[:, ny, nx, tr, tc, :]
    assert (yy - y).abs().max() < 1e-12
    # backward: D[s][tap, co] = sum over the children p of s of dc[p - tap + 1]  (a 4x4 window of dc around the children)
    dc = torch.randn_like(y)
    (y * dc).sum().backward()
    dcp = F.pad(dc, (1, 1, 1, 1))
    D = torch.zeros(n, hs, ws, 3, 3, co, dtype=DT)
    dWs = torch.zeros(co, cs, 3, 3, dtype=DT)
    for sy in range(hs):
        for sx in range(ws):
            win = dcp[:, :, 2 * sy:2 * sy + 4, 2 * sx:2 * sx + 4]          # win[i][j] <-> dc[2 sy - 1 + i][2 sx - 1 + j]
            for tr in range(3):
                for tc in range(3):
                    D[:, sy, sx, tr, tc, :] = win[:, :, 2 - tr, 2 - tc] + win[:, :, 2 - tr, 3 - tc] + win[:, :, 3 - tr, 2 - tc] + win[:, :, 3 - tr, 3 - tc]
                    for a in range(2):
                        for b in range(2):
                            dWs[:, :, tr, tc] += torch.einsum("no,nc->oc", win[:, :, a - tr + 2, b - tc + 2], v[:, :, 2 * sy + a, 2 * sx + b])
    assert (torch.einsum("nhwtso,octs->nchw", D, W[:, :cu]) - u_.grad).abs().max() < 1e-12        # gradient AT SOURCE resolution
    assert (torch.einsum("nhwtso,nchw->octs", D, u) - W_.grad[:, :cu]).abs().max() < 1e-12
    assert (dWs - W_.grad[:, cu:]).abs().max() < 1e-12


def _space_to_depth(x):
    n, c, h, w = x.shape
    return x.view(n, c, h // 2, 2, w // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, h // 2, w // 2)   # channel = (a, b, ch)


def test_space_to_depth_stem_identity():
    torch.manual_seed(1)
    n, c, co, h, w = 2, 3, 5, 12, 16
    x = torch.randn(n, c, h, w, dtype=DT)
    W = torch.randn(co, c, 7, 7, dtype=DT, requires_grad=True)
    y = F.conv2d(x, W, stride=2, padding=3)
    # wsub[co][(a, b, ch)][ja][jb] = w[co][ch][2 ja + a - 1][2 jb + b - 1]   (taps -1 and 7 do not exist: zero)
    Ws = torch.zeros(co, 4 * c, 4, 4, dtype=DT)
    for ja in range(4):
        for jb in range(4):
            for a in range(2):
                for b in range(2):
                    ty, tx = 2 * ja + a - 1, 2 * jb + b - 1
                    if 0 <= ty < 7 and 0 <= tx < 7:
                        Ws[:, (a * 2 + b) * c:(a * 2 + b + 1) * c, ja, jb] = W.detach()[:, :, ty, tx]
    xs = F.pad(_space_to_depth(x), (2, 1, 2, 1))                 # 2 cells in front; the symmetric padding's extra output is never computed
    ys = F.conv2d(xs, Ws)
    assert ys.shape == y.shape and (ys - y).abs().max() < 1e-12
    # weight gradient: the 4x4 problem's gradient, gathered back: dw[ty][tx] = dwsub[(ty+1)>>1][(tx+1)>>1][((ty+1)&1, (tx+1)&1)]
    dc = torch.randn_like(y)
    (y * dc).sum().backward()
    Ws_ = Ws.clone().requires_grad_(True)
    (F.conv2d(xs, Ws_) * dc).sum().backward()
    g = torch.zeros_like(W)
    for ty in range(7):
        for tx in range(7):
            a, b = (ty + 1) & 1, (tx + 1) & 1
            g[:, :, ty, tx] = Ws_.grad[:, (a * 2 + b) * c:(a * 2 + b + 1) * c, (ty + 1) >> 1, (tx + 1) >> 1]
    assert (g - W.grad).abs().max() < 1e-10


def _sp_axis(k, d, p):
    """effective taps of one axis per output parity q: offsets e0 .. e0+ne-1 into the SOURCE, and which original taps land on each"""
    out = []
    for q in range(2):
        offs = [(q + t * d - p) // 2 for t in range(k)]         # floor division: source offset read by tap t
        e0 = min(offs)
        groups = [[t for t in range(k) if offs[t] == e0 + e] for e in range(max(offs) - e0 + 1)]
        out.append((e0, groups))
    return out


def test_sub_pixel_decomposition_of_conv_over_upsampled_source():
    torch.manual_seed(2)
    for k, d, p in ((3, 1, 1), (3, 2, 2), (1, 1, 0), (4, 1, 2)):
        n, c, co, hs, ws = 1, 4, 3, 6, 7
        u = torch.randn(n, c, hs, ws, dtype=DT)
        W = torch.randn(co, c, k, k, dtype=DT)
        y = F.conv2d(F.interpolate(u, scale_factor=2, mode="nearest"), W, padding=p, dilation=d)
        if y.shape[-2:] != (2 * hs, 2 * ws):
            y = y[:, :, :2 * hs, :2 * ws]
        ay, ax = _sp_axis(k, d, p), _sp_axis(k, d, p)
        for py in range(2):
            for px in range(2):
                (ey0, gy), (ex0, gx) = ay[py], ax[px]
                Weff = torch.zeros(co, c, len(gy), len(gx), dtype=DT)
                for i, ty in enumerate(gy):
                    for j, tx in enumerate(gx):
                        for a in ty:
                            for b in tx:
                                Weff[:, :, i, j] += W[:, :, a, b]
                # class output (ky, kx) reads source (ky + ey0 + i, kx + ex0 + j): pad the source so that every index exists
                lo_y, lo_x = max(0, -ey0), max(0, -ex0)
                hi_y, hi_x = max(0, ey0 + len(gy) - 1), max(0, ex0 + len(gx) - 1)
                up = F.pad(u, (lo_x, hi_x, lo_y, hi_y))
                yc = F.conv2d(up, Weff)[:, :, ey0 + lo_y:ey0 + lo_y + hs, ex0 + lo_x:ex0 + lo_x + ws]
                ref = y[:, :, py::2, px::2]
                assert yc.shape == ref.shape and (yc - ref).abs().max() < 1e-12, (k, d, p, py, px)
