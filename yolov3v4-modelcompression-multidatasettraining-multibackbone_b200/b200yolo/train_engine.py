"""Training-mode executor: forward with batch-statistics BatchNorm and the full backward pass on sm_100a.

Replaces autograd over ~500 ATen/cuDNN calls (reference train.py:380-441 -> models.py:508-551) by a static plan:

  forward   conv (tcgen05, per-channel sum / sum^2 from the epilogue) -> bn_finalize -> bn+act(+shortcut) apply
  backward  yolo permute^T -> [bn+act backward reduce / apply -> wgrad (tcgen05, pixel-K GEMM) -> dgrad (tcgen05)]

Activations z (raw conv output) and y (post activation) are kept in NHWC fp16 (accurate forward).  Gradients w.r.t.
activations (dY) are stored in **bf16** (fp32 range: their magnitude drifts by orders of magnitude across ~100
BatchNorm layers, which fp16 cannot hold with one loss scale).  The conv data gradient dZ, which is an operand of the
two tensor-core GEMMs (dgrad with the fp16 weights, wgrad with the fp16 activations), is written in fp16 times a
per-layer power-of-two scale chosen inside the BN-backward kernel from a bound on max|dZ| (tcgen05 kind::f16 faults on
mixed f16/bf16 operands); the GEMM epilogues read 1/scale from device memory, so there is no host synchronisation.
Weight / BN gradients are accumulated and stored in fp32.
The plan plugs into autograd as ONE torch.autograd.Function whose inputs are the model parameters, so
loss.backward(), DistributedDataParallel hooks and torch optimisers work unchanged.
"""
import ctypes as C

import os

import torch
import torch.nn as nn

from . import ops
from .engine import LazyFeatures, _Tensor, _block_parts
from .lib import ACT, ConvDesc, OUT_F16, call, ptr, stream_ptr

GDT = torch.bfloat16   # gradient tensors: bf16 (range of fp32) -- fp16 over/underflows across ~100 BN layers
HEAD_PAD = 256  # head convs (255 filters) are run with 256 output channels (zero row) so that K % 16 == 0 in dgrad


def _round_up(v, m):
    return (v + m - 1) // m * m


def _widen(view, channels):
    """The same NHWC view with `channels` channels (its buffer is at least that wide and zero beyond the view)."""
    B, H, W, C = view.shape
    return torch.as_strided(view, (B, H, W, channels), view.stride(), view.storage_offset())


class _ConvRec:
    __slots__ = ('i', 'conv', 'bn', 'act', 'slope', 'src', 'z', 'y', 'res', 'head', 'stem', 'k', 's', 'p', 'Cout',
                 'Cpad', 'w16', 'wT', 'stats', 'mean', 'invstd', 'scale', 'shift', 'ones', 'zeros', 'w32', 'aux_row', 'wstem',
                 'save', 'sums', 'dw', 'fast_bn', 'depthwise', 'Ipad', 'dwg')

    def __init__(self):
        for k in self.__slots__:
            setattr(self, k, None)


class TrainPlan:
    def __init__(self, model, x_shape, device, keep_features=False):
        self.model = model
        self.device = device
        self.keep_features = bool(keep_features)
        self.B, self.Cin, self.H, self.W = x_shape
        self.convs = []
        self.order = []        # forward step list
        self.param_version = None
        self._build()

    # ---------------------------------------------------------------------------------------------------------
    def _build(self):
        model, B, dev = self.model, self.B, self.device
        defs, mods, routs = model.module_defs, model.module_list, model.routs
        n = len(defs)
        shapes, prev = [], (self.Cin, self.H, self.W)
        for i, (d, m) in enumerate(zip(defs, mods)):
            t = d['type']
            Cc, H, W = prev
            if t in ('convolutional', 'depthwise'):
                conv = _block_parts(m)[0]
                H, W = ops.conv_out_hw(H, W, conv.kernel_size[0], conv.stride[0], conv.padding[0])
                Cc = conv.out_channels
            elif t == 'se':
                pass
            elif t == 'maxpool':
                k, s = d['size'], d['stride']
                if k == 2 and s == 1:
                    H, W = (H + 1 - k) // s + 1, (W + 1 - k) // s + 1
                else:
                    pd = (k - 1) // 2
                    H, W = (H + 2 * pd - k) // s + 1, (W + 2 * pd - k) // s + 1
            elif t == 'upsample':
                H, W = H * d['stride'], W * d['stride']
            elif t == 'route':
                srcs = [i + l if l < 0 else l for l in d['layers']]
                Cc = sum(shapes[s][0] for s in srcs)
                if 'groups' in d:
                    Cc //= 2
                H, W = shapes[srcs[0]][1], shapes[srcs[0]][2]
            elif t in ('shortcut', 'yolo'):
                pass
            else:
                raise NotImplementedError("layer type '%s' is not supported by the sm_100a training engine yet" % t)
            prev = (Cc, H, W)
            shapes.append(prev)

        is_head = [False] * n
        fused_into = [None] * n   # shortcut i folded into the bn+act apply of conv i-1
        for i, d in enumerate(defs):
            if d['type'] == 'yolo' and i > 0 and defs[i - 1]['type'] == 'convolutional':
                is_head[i - 1] = True
            if d['type'] == 'shortcut' and i > 1:
                frm = [i + l if l < 0 else l for l in d['from']]
                if (not self.keep_features
                        and len(frm) == 1 and defs[i - 1]['type'] == 'convolutional' and not routs[i - 1]
                        and not getattr(mods[i], 'weight', False) and shapes[frm[0]] == shapes[i - 1]
                        and _block_parts(mods[i - 1])[1] is not None):
                    fused_into[i] = i - 1

        tens = [None] * n
        for i, d in enumerate(defs):
            t = d['type']
            if t == 'yolo' or (t == 'route' and len(d['layers']) == 1 and 'groups' not in d):
                continue
            Cc, H, W = shapes[i]
            tens[i] = _Tensor(Cc, H, W, torch.float32 if is_head[i] else torch.float16)
        self.pre_add = {}
        for i in range(n):
            if fused_into[i] is not None:
                tens[fused_into[i]] = tens[i]
        for i, d in enumerate(defs):
            if d['type'] == 'route' and len(d['layers']) == 1 and 'groups' not in d:
                l = d['layers'][0]
                src = i + l if l < 0 else l
                tens[i] = tens[src]

        # concat placement (same policy as the inference plan); gradients mirror the placement
        placed, copies = set(), {}
        self.bufs = []   # (activation buffer, gradient buffer) pairs to keep alive

        def new_buf(Cc, H, W, dtype=torch.float16):
            if dtype != torch.float16:
                return torch.empty((B, H, W, _round_up(Cc, 4)), dtype=dtype, device=dev)
            if Cc % 16 == 0:
                return torch.empty((B, H, W, Cc), dtype=dtype, device=dev)
            # MobileNet widths (24, 40, 72 ...): pitch rounded up to the MMA K granule, pad channels stay zero forever
            return torch.zeros((B, H, W, _round_up(Cc, 16)), dtype=dtype, device=dev)

        def zeros_later(shape, dtype):
            """placeholder (meta tensor) for a buffer that is carved out of the zero-at-start-of-backward arena"""
            return torch.empty(shape, dtype=dtype, device='meta')

        self.grad_of = {}  # id(_Tensor) -> gradient _Tensor (fp16, same placement)

        def make_grad(t, gbuf=None, c0=0):
            g = _Tensor(t.C, t.H, t.W, GDT)
            g.buf, g.c0 = gbuf, c0
            self.grad_of[id(t)] = g
            return g

        for i, d in enumerate(defs):
            if d['type'] != 'route' or len(d['layers']) == 1:
                continue
            srcs = [i + l if l < 0 else l for l in d['layers']]
            dst = tens[i]
            dst.buf = new_buf(dst.C, dst.H, dst.W)
            gdst = make_grad(dst, zeros_later(dst.buf.shape, GDT))
            off = 0
            for s in srcs:
                st = tens[s]
                if id(st) not in placed and st.buf is None and off % 8 == 0 and st.dtype == torch.float16 \
                        and st is not dst:
                    st.buf, st.c0 = dst.buf, off
                    make_grad(st, gdst.buf, off)
                    placed.add(id(st))
                else:
                    copies.setdefault(i, []).append((st, off))
                off += st.C

        def alloc(t):
            if t.buf is None:
                t.buf = new_buf(t.C if t.dtype == torch.float16 else HEAD_PAD, t.H, t.W, t.dtype)
                t.c0 = 0
            if id(t) not in self.grad_of:
                if t.dtype == torch.float16:
                    make_grad(t, zeros_later(t.buf.shape, GDT))
                else:  # head output: its gradient is directly the GEMM operand dZ -> fp16 (times a device-side scale)
                    g = _Tensor(HEAD_PAD, t.H, t.W, torch.float16)
                    g.buf = zeros_later((B, t.H, t.W, HEAD_PAD), torch.float16)
                    self.grad_of[id(t)] = g
            return t

        self.yolo = []
        self.feature_views = []
        feat_idx = {}
        for i, (d, m) in enumerate(zip(defs, mods)):
            t = d['type']
            if t == 'se':
                fc = m[0].fc if isinstance(m, nn.Sequential) else m.fc
                src = tens[i - 1]
                out = alloc(tens[i])
                cr = fc[0].out_features
                rec = {'i': i, 'src': src, 'y': out, 'fc1': fc[0], 'fc2': fc[2],
                       'ws': ops.se_workspace(B, src.C, cr, dev), 'ws_bwd': ops.se_workspace(B, src.C, cr, dev)}
                self.order.append(('se', rec))
                feat_idx[i] = out
            elif t in ('convolutional', 'depthwise'):
                conv, bn, act, slope = _block_parts(m)
                r = _ConvRec()
                r.depthwise = t == 'depthwise'
                if r.depthwise:
                    if conv.groups != conv.in_channels or conv.out_channels != conv.in_channels or bn is None or i == 0:
                        raise NotImplementedError("depthwise block: channel multiplier 1 with BatchNorm expected")
                elif conv.groups != 1:
                    raise NotImplementedError("grouped convolution is not supported by the sm_100a engine yet")
                r.i, r.conv, r.bn, r.act, r.slope = i, conv, bn, act, slope
                r.k, r.s, r.p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
                r.head, r.stem = is_head[i], (i == 0)
                if r.stem and conv.in_channels > 4:
                    raise NotImplementedError("first layer with more than 4 input channels")
                r.Cout = conv.out_channels
                # GEMM K granule: Cout is the K of the data-gradient GEMM, Cin the K of the forward / weight-gradient one
                r.Cpad = HEAD_PAD if r.head else _round_up(r.Cout, 16)
                r.Ipad = _round_up(conv.in_channels, 16)
                if r.head and (r.Cout > HEAD_PAD or bn is not None or act != 'linear'):
                    raise NotImplementedError("unsupported YOLO head conv")
                r.src = None if i == 0 else tens[i - 1]
                r.y = alloc(tens[i])
                if i + 1 < n and fused_into[i + 1] == i:
                    l = defs[i + 1]['from'][0]
                    r.res = tens[i + 1 + l if l < 0 else l]
                if bn is not None:
                    r.z = _Tensor(r.Cout, r.y.H, r.y.W)
                    r.z.buf = new_buf(r.Cout, r.y.H, r.y.W)
                    r.stats = zeros_later((2, r.Cout), torch.float32)
                    r.sums = zeros_later((2, r.Cout), torch.float32)
                    r.fast_bn = r.Cout % 8 == 0 and r.Cout // 8 <= 256
                else:
                    r.z = r.y   # no BN: the conv epilogue applies bias + activation directly
                    r.ones = torch.ones(r.Cpad, dtype=torch.float32, device=dev)
                    r.zeros = torch.zeros(r.Cpad, dtype=torch.float32, device=dev)
                r.aux_row = len(self.convs)
                self.convs.append(r)
                self.order.append(('conv', r))
                if not is_head[i] and r.res is None:
                    feat_idx[i] = r.y   # (a conv with a fused shortcut never materialises its pre-add output)
                if r.depthwise:
                    r.dwg = zeros_later(tuple(conv.weight.shape), torch.float32)
            elif t == 'shortcut':
                if fused_into[i] is not None:
                    continue
                frm = [i + l if l < 0 else l for l in d['from']]
                if getattr(m, 'weight', False):
                    raise NotImplementedError("weighted shortcut is not supported yet")
                out = alloc(tens[i])
                self.order.append(('add', tens[i - 1], [tens[s] for s in frm], out))
            elif t == 'route':
                if len(d['layers']) == 1:
                    if 'groups' in d:
                        src, g = tens[i - 1], tens[i]
                        g.buf, g.c0 = src.buf, src.c0 + src.C // 2
                        gs = self.grad_of[id(src)]
                        make_grad(g, gs.buf, gs.c0 + src.C // 2)
                    continue
                for st, off in copies.get(i, []):
                    self.order.append(('copy', st, tens[i], off))
            elif t == 'upsample':
                self.order.append(('upsample', tens[i - 1], alloc(tens[i]), d['stride']))
            elif t == 'maxpool':
                k, s = d['size'], d['stride']
                self.order.append(('maxpool', tens[i - 1], alloc(tens[i]), k, s, k == 2 and s == 1))
                if m.__class__.__name__ == 'Sequential':
                    feat_idx[i] = tens[i]
            elif t == 'yolo':
                self.yolo.append((m, tens[i - 1]))
        for i, m in enumerate(mods):
            if m.__class__.__name__ == 'Sequential' and i + 1 < n and defs[i + 1]['type'] != 'yolo':
                self.feature_views.append(feat_idx.get(i))
        self.anchors_px = [m.anchors.to(dev).float().contiguous() for (m, _) in self.yolo]
        self.bn_counters = [r.bn.num_batches_tracked for r in self.convs if r.bn is not None]
        self._plan_grad_writes()
        # ---- arenas ------------------------------------------------------------------------------------------------
        # Everything that has to be zero when a pass starts is carved out of ONE allocation per pass, cleared by ONE
        # memset (round 1: one fill launch per buffer, ~450 launches per step):
        #   forward arena : the BatchNorm channel sums (sum z, sum z^2) the conv epilogues accumulate into
        #   backward arena: activation gradients (bf16; consumers accumulate into them), BN backward sums, the per-layer
        #                   [max|du|, s, 1/s] scale rows, and the packed fp32 weight gradients (split-K red.add targets)
        self.dz_aux = torch.empty((len(self.convs) + 1, 4), dtype=torch.float32, device='meta')
        for r in self.convs:
            if not r.stem and not r.depthwise:
                r.dw = torch.empty((r.Cpad, r.k, r.k, r.Ipad), dtype=torch.float32, device='meta')
        fwd_req, bwd_req, seen = [], [], {}

        def want(lst, holder, attr):
            t = getattr(holder, attr)
            if t is None or t.device.type != 'meta':
                return
            key = id(t)
            if key not in seen:
                seen[key] = []
                lst.append(t)
            seen[key].append((holder, attr))

        for r in self.convs:
            want(fwd_req, r, 'stats')
        for g in self.grad_of.values():
            want(bwd_req, g, 'buf')
        for r in self.convs:
            want(bwd_req, r, 'sums')
            want(bwd_req, r, 'dw')
            want(bwd_req, r, 'dwg')
        want(bwd_req, self, 'dz_aux')

        # gradient buffers whose first writer overwrites (see _plan_grad_writes) go behind the part that is memset
        grad_ids = {id(g.buf) for g in self.grad_of.values()}
        is_clean = lambda t: id(t) in grad_ids and id(t) not in self.dirty_grad_bufs
        bwd_req = [t for t in bwd_req if not is_clean(t)] + [t for t in bwd_req if is_clean(t)]
        self.bwd_zero_bytes = sum(_round_up(t.numel() * t.element_size(), 256) for t in bwd_req if not is_clean(t))

        def carve(reqs):
            offs, total = [], 0
            for t in reqs:
                offs.append(total)
                total += _round_up(t.numel() * t.element_size(), 256)
            arena = torch.zeros(max(total, 256), dtype=torch.uint8, device=dev)
            for t, off in zip(reqs, offs):
                nbytes = t.numel() * t.element_size()
                real = arena[off:off + nbytes].view(t.dtype).view(t.shape)
                for holder, attr in seen[id(t)]:
                    setattr(holder, attr, real)
            return arena

        self.fwd_arena = carve(fwd_req)
        self.bwd_arena = carve(bwd_req)
        # BatchNorm statistics saved by the forward for the backward: [mean, invstd, scale, shift] per layer
        for r in self.convs:
            if r.bn is not None:
                r.save = torch.empty((4, r.Cout), dtype=torch.float32, device=dev)
        # The weight gradients hang off the critical path (dz_i -> dW_i has no consumer until the optimiser), so they run
        # on a second stream concurrently with the data-gradient / BatchNorm chain.  dz is therefore kept per layer
        # (no buffer shared between a running wgrad and the next layer's BN backward); B2Y_WGRAD_STREAM=0 serialises.
        self.side_wgrad = os.environ.get('B2Y_WGRAD_STREAM', '1') != '0'
        self.dz_bufs = {}
        if not self.side_wgrad:
            maxz = max(r.Cpad * r.y.H * r.y.W for r in self.convs)
            self.dz_scratch = torch.empty(B * maxz, dtype=torch.float16, device=dev)
        self.side_stream = None
        maxc = max(r.Cpad for r in self.convs)
        self.dgb_scratch = torch.empty((2, maxc), dtype=torch.float32, device=dev)
        self.stem_dw = torch.empty(maxc * 32, dtype=torch.float32, device=dev)
        self.params = [p for p in model.parameters()]
        # plan-owned gradient outputs for the plain-autograd path (no flat sink): OIHW fp32 per parameter
        self.own_grads = None
        self.pack_table = self.unpack_table = None
        self.pack_key = self.unpack_key = None
        self.runs = 0
        self.fwd_graph = self.bwd_graph = None
        self.bwd_key = None

    # ---------------------------------------------------------------------------------------------------------
    def _params_version(self):
        return sum(p._version for p in self.params)

    def _pack(self):
        """fp32 master weights -> fp16 operand layouts (forward [O][k][k][I], per-phase data-gradient slabs): ONE
        table-driven launch over all tcgen05 convolutions (csrc/multi.cu) + the stem's own tiny pack."""
        from .lib import PackItem, raw
        body = [r for r in self.convs if not r.stem and not r.depthwise]
        key = tuple(r.conv.weight.data_ptr() for r in body)
        if self.pack_table is None or key != self.pack_key:
            items, tiles = [], 0
            for r in body:
                O, I, k = r.conv.weight.shape[0], r.conv.weight.shape[1], r.k
                if r.w16 is None:
                    # zero-filled once: the K-pad columns (>= I) and pad rows (>= O) are never written again
                    r.w16 = torch.zeros((r.Cpad, k, k, r.Ipad), dtype=torch.float16, device=self.device)
                    r.wT = torch.zeros(r.Cpad * I * k * k, dtype=torch.float16, device=self.device)
                w = r.conv.weight.detach()
                assert w.is_contiguous() and w.dtype == torch.float32
                ti = int(raw().b2y_layout_tile_i(k))
                if ti <= 0:
                    raise NotImplementedError("kernel size %d is not supported by the training engine" % k)
                items.append(PackItem(w.data_ptr(), r.w16.data_ptr(), r.wT.data_ptr(), O, r.Cpad, I, k, r.s, r.p,
                                      tiles, r.Ipad))
                tiles += ((r.Cpad + 31) // 32) * ((I + ti - 1) // ti)
            arr = (PackItem * len(items))(*items)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self.pack_table = (host.to(self.device), len(items), tiles)
            self.pack_key = key
        tab, n, tiles = self.pack_table
        call("b2y_pack_conv_weights_multi", ptr(tab), n, tiles, stream_ptr())
        for r in self.convs:
            if r.stem:
                w = r.conv.weight.detach()
                if r.head:
                    pad = torch.zeros((HEAD_PAD - r.Cout,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
                    w = torch.cat([w, pad], 0)
                r.w32 = w.float().contiguous()
                # tensor-core stem: full-im2col weight layout when the receptive field fits one 64-byte GEMM row
                r.wstem = ops.pack_stem_weights(r.w32) if w.shape[1] * r.k * r.k <= 32 and not r.head else None

    def _plan_grad_writes(self):
        """Which backward write into an activation-gradient buffer is the FIRST one of its channel range.

        Consumers used to accumulate into gradient buffers that the start-of-backward memset had cleared: for yolov4 at
        8 x 640^2 that is ~2.4 GB of memset per step plus a read of zeros by every data-gradient GEMM (another ~2 GB).
        The launch order of the backward pass is static, so the first writer of every (buffer, channel range) is known
        when the plan is built: it OVERWRITES (dgrad accumulate=False, copy instead of add) and the buffer leaves the
        zeroed arena.  A buffer stays 'dirty' (zeroed, accumulate-only) when a first writer cannot overwrite (up-sample /
        max-pool / depthwise / SE backward, stride > kernel dgrad), when ranges overlap partially, or when something would
        read channels nobody has written.  B2Y_GRAD_FIRSTWRITE=0 keeps every buffer dirty."""
        self.first_write, cov, dirty = {}, {}, set()
        G = lambda t: self.grad_of[id(t)]
        enabled = os.environ.get('B2Y_GRAD_FIRSTWRITE', '1') != '0'

        def write(g, n, key, can_overwrite):
            b, R = id(g.buf), set(range(g.c0, g.c0 + n))
            c = cov.setdefault(b, set())
            inter = R & c
            if not inter and can_overwrite and enabled:
                self.first_write[key] = True
            else:
                self.first_write[key] = False
                if inter != R:              # partial overlap, or a first writer that can only accumulate
                    dirty.add(b)
            c |= R

        def read(g, n, c0=None):
            b = id(g.buf)
            c0 = g.c0 if c0 is None else c0
            if not set(range(c0, c0 + n)) <= cov.get(b, set()):
                dirty.add(b)

        for (m, raw) in self.yolo:              # head gradients: written by yolo_grad_to_raw only where a loss exists
            g = G(raw)
            dirty.add(id(g.buf))
            cov[id(g.buf)] = set(range(0, HEAD_PAD))
        for idx in range(len(self.order) - 1, -1, -1):
            st = self.order[idx]
            kind = st[0]
            if kind == 'conv':
                r = st[1]
                read(G(r.y), G(r.y).C)
                if r.bn is not None and r.res is not None:
                    write(G(r.res), G(r.res).C, ('res', r.i), True)
                if r.stem:
                    continue
                if r.depthwise:
                    write(G(r.src), r.conv.in_channels, ('dwd', r.i), False)
                else:
                    write(G(r.src), r.conv.in_channels, ('dgrad', r.i), r.k >= r.s)
            elif kind == 'se':
                e = st[1]
                read(G(e['y']), G(e['y']).C)
                write(G(e['src']), G(e['src']).C, ('se', idx), False)
            elif kind == 'add':
                _, first, others, out = st
                go = G(out)
                read(go, go.C)
                for j, t in enumerate([first] + list(others)):
                    write(G(t), min(G(t).C, go.C), ('add', idx, j), True)
            elif kind == 'copy':
                _, srct, dst, off = st
                gd = G(dst)
                read(gd, srct.C, gd.c0 + off)
                write(G(srct), srct.C, ('copy', idx), True)
            elif kind in ('upsample', 'maxpool'):
                src, out = st[1], st[2]
                read(G(out), G(out).C)
                write(G(src), G(src).C, (kind, idx), False)
        self.dirty_grad_bufs = dirty

    def _grad_dst(self, param):
        """Where the gradient of `param` is written: its slice of the flat data-parallel buffer (sink), else a
        plan-owned fp32 tensor (handed to autograd as a copy)."""
        if self.sink is not None:
            d = self.sink.get(id(param))
            if d is not None:
                return d
        if self.own_grads is None:
            self.own_grads = {}
        d = self.own_grads.get(id(param))
        if d is None:
            d = torch.zeros_like(param, dtype=torch.float32, memory_format=torch.contiguous_format)
            self.own_grads[id(param)] = d
        return d

    def _unpack_all(self):
        """packed fp32 weight gradients [O][k][k][I] of all tcgen05 convolutions -> OIHW destinations, ONE launch."""
        from .lib import UnpackItem, raw
        body = [r for r in self.convs if not r.stem and not r.depthwise]
        dsts = [self._grad_dst(r.conv.weight) for r in body]
        key = tuple(d.data_ptr() for d in dsts)
        if self.unpack_table is None or key != self.unpack_key:
            items, tiles = [], 0
            for r, d in zip(body, dsts):
                O, I, k = r.Cout, r.conv.in_channels, r.k
                ti = int(raw().b2y_layout_tile_i(k))
                items.append(UnpackItem(r.dw.data_ptr(), d.data_ptr(), O, I, k, 0, tiles, r.Ipad))
                tiles += ((O + 31) // 32) * ((I + ti - 1) // ti)
            arr = (UnpackItem * len(items))(*items)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self.unpack_table = (host.to(self.device), len(items), tiles)
            self.unpack_key = key
        tab, n, tiles = self.unpack_table
        call("b2y_unpack_wgrad_multi", ptr(tab), n, tiles, stream_ptr())

    def _zview(self, r):
        return r.z.buf[..., r.z.c0:r.z.c0 + r.Cpad] if r.head else r.z.view()

    # ---- CUDA-graph replay of the two static launch sequences (forward, backward) ------------------------------
    # A training step is ~15 launches per conv layer; issued eagerly from Python the step is CPU-launch bound at 8
    # images per GPU.  Shapes, buffers and the master parameters are static, the per-layer gradient scales live on the
    # device, and the (dynamic) target list only enters the loss between the two graphs.
    def _graphs_enabled(self):
        return getattr(self.model, 'use_cuda_graph', os.environ.get('B2Y_NO_GRAPH', '0') != '1')

    def forward_graphed(self, x):
        if not self._graphs_enabled():
            return self.forward(x)
        if self.runs < 1:                      # first step eager: lazy allocations, warm caches
            self.runs += 1
            return self.forward(x)
        if self.fwd_graph is None:
            self.static_x = torch.empty((self.B, self.Cin, self.H, self.W), dtype=torch.float32, device=self.device)
            self.static_x.copy_(x)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.static_outs = self.forward(self.static_x, force_pack=True)
            self.fwd_graph = g
            self.bwd_graph = None
        self.static_x.copy_(x)
        self.fwd_graph.replay()
        return [o.view_as(o) for o in self.static_outs]

    def backward_graphed(self, dps):
        if not self._graphs_enabled() or self.fwd_graph is None:
            # plan-owned gradient buffers are overwritten by the next step: autograd gets copies
            return {p: g.clone() for p, g in self.backward(dps).items()}
        sink = getattr(self.model, '_b2y_grad_sink', None)
        key = (id(sink), float(getattr(self.model, 'grad_scale', None) or 1.0))
        if self.bwd_graph is None or key != self.bwd_key:
            self.static_dps = [torch.zeros_like(o, dtype=torch.float32) for o in self.static_outs]
            for sd, d in zip(self.static_dps, dps):
                if d is not None:
                    sd.copy_(d)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.fwd_graph.pool()):
                self.static_grads = self.backward(self.static_dps)
            self.bwd_graph, self.bwd_key = g, key
        for sd, d in zip(self.static_dps, dps):
            if d is None:
                sd.zero_()
            else:
                sd.copy_(d)
        self.bwd_graph.replay()
        # gradients that are not written into a flat sink live in graph-owned memory: hand out copies
        return {p: g.clone() for p, g in self.static_grads.items()}

    def forward(self, x, force_pack=False):
        model = self.model
        ver = self._params_version()
        if force_pack or ver != self.param_version:
            self._pack()
            self.param_version = None if force_pack else ver
        x = x.contiguous().float()
        self.x = x
        self.fwd_arena.zero_()                     # all BatchNorm channel sums (one memset)
        for st in self.order:
            kind = st[0]
            if kind == 'conv':
                r = st[1]
                self._conv_forward(r, x)
            elif kind == 'se':
                e = st[1]
                ops.se_fwd(e['src'].view(), e['fc1'].weight.detach(), e['fc2'].weight.detach(), out=e['y'].view(),
                           ws=e['ws'])
            elif kind == 'add':
                cur = st[1]
                for s in st[2]:                     # channel-sliced when the widths differ (layers.py:57-72)
                    xv, av, ov = cur.view(), s.view(), st[3].view()
                    n = min(xv.shape[3], av.shape[3])
                    ops.add(xv[..., :n], av[..., :n], out=ov[..., :n])
                    if xv.shape[3] > n and cur is not st[3]:
                        ops.copy_channels(xv[..., n:], ov[..., n:])
                    cur = st[3]
            elif kind == 'copy':
                _, srct, dst, off = st
                ops.copy_channels(srct.view(), dst.buf[..., off:off + srct.C])
            elif kind == 'upsample':
                ops.upsample(st[1].view(), st[3], out=st[2].view())
            elif kind == 'maxpool':
                ops.maxpool(st[1].view(), st[3], st[4], tiny_pad=st[5], out=st[2].view())
        if self.bn_counters:
            torch._foreach_add_(self.bn_counters, 1)

        outs = []
        for (m, raw), anc in zip(self.yolo, self.anchors_px):
            _, p = ops.yolo_decode(raw.buf, m.na, m.no, anc, m.stride, io=None)
            m.nx, m.ny = raw.W, raw.H
            outs.append(p)
        return outs

    def _conv_forward(self, r, x):
        if r.bn is None:
            bias = r.conv.bias.detach() if r.conv.bias is not None else None
            if r.head and bias is not None:
                bias = torch.cat([bias, bias.new_zeros(HEAD_PAD - r.Cout)])
            out = self._zview(r)
            if r.stem:
                ops.stem_conv(x, r.w32, bias, r.k, r.s, r.p, act=r.act, slope=r.slope, out=out)
            else:
                ops.conv2d(r.src.view_k(), r.w16, bias, r.k, r.s, r.p, act=r.act, slope=r.slope, out=out)
            return
        z = r.z.view()
        if r.stem and getattr(r, 'wstem', None) is not None:
            # im2col workspace [B*Ho*Wo][32] fp16 (kept for the weight gradient) + one-k-step GEMM with the BN sums
            # coming out of its epilogue
            d = ops.make_conv_desc((self.B, self.H, self.W, r.conv.in_channels), r.conv.in_channels, r.Cout, r.k, r.s,
                                   r.p, r.Cout)
            if getattr(self, 'stem_ws', None) is None:
                self.stem_ws = ops.stem_workspace(d, self.device)
            ops.stem_conv_tc(x, r.wstem, None, r.conv.in_channels, r.k, r.s, r.p, act='linear', out=z,
                             workspace=self.stem_ws, stats=(r.stats[0], r.stats[1]))
        elif r.stem:
            ops.stem_conv(x, r.w32, None, r.k, r.s, r.p, act='linear', out=z)
            # channel sums of the stem output: the reduce kernel with u = z, dy = z gives (sum z, sum z*z)
            if r.ones is None:
                r.ones = torch.ones(r.Cout, dtype=torch.float32, device=self.device)
                r.zeros = torch.zeros(r.Cout, dtype=torch.float32, device=self.device)
            B_, H_, W_, C_ = z.shape
            call("b2y_bn_act_bwd_reduce", ptr(z), ops._pitch(z), ptr(z), ops._pitch(z), ptr(r.ones), ptr(r.zeros),
                 ptr(r.zeros), ptr(r.ones), ptr(r.stats[1]), ptr(r.stats[0]), None, B_ * H_ * W_, C_, 0, 0.0, 0,
                 stream_ptr())
        elif r.depthwise:
            ops.dwconv2d(r.src.view(), r.conv.weight.detach(), r.k, r.s, r.p, out=z, stats=(r.stats[0], r.stats[1]))
        else:
            ops.conv2d(r.src.view_k(), r.w16[:r.Cout], None, r.k, r.s, r.p, out=z, stats=(r.stats[0], r.stats[1]))
        bn = r.bn
        count = z.shape[0] * z.shape[1] * z.shape[2]
        res = r.res.view() if r.res is not None else None
        if r.fast_bn:
            y = r.y.view()
            call("b2y_bn_train_fwd", ptr(z), ops._pitch(z), ptr(r.stats[0]), ptr(r.stats[1]), int(count),
                 ptr(bn.weight.detach()), ptr(bn.bias.detach()), float(bn.eps), float(bn.momentum),
                 ptr(bn.running_mean), ptr(bn.running_var), ptr(r.save), ptr(res),
                 ops._pitch(res) if res is not None else 0, ptr(y), ops._pitch(y), int(count), r.Cout,
                 ACT[r.act], float(r.slope), stream_ptr())
            r.mean, r.invstd, r.scale, r.shift = r.save[0], r.save[1], r.save[2], r.save[3]
            return
        r.mean, r.invstd, r.scale, r.shift = ops.bn_finalize(r.stats[0], r.stats[1], count, bn.weight.detach(),
                                                             bn.bias.detach(), bn.eps, bn.momentum, bn.running_mean,
                                                             bn.running_var)
        ops.bn_act_fwd(z, r.scale, r.shift, r.act, r.slope, residual=res, out=r.y.view())

    # ---------------------------------------------------------------------------------------------------------
    def backward(self, dps):
        """dps: gradients w.r.t. the yolo outputs p (fp32). Returns {param: grad} for every model parameter."""
        # gradients flow in bf16, so no loss scaling is needed; model.grad_scale (power of two) is still honoured
        S = float(getattr(self.model, 'grad_scale', None) or 1.0)
        self.last_grad_scale = S
        inv = 1.0 / S
        self.sink = getattr(self.model, '_b2y_grad_sink', None)   # FlatDataParallel: write into the flat buffer
        # one memset: BN sums, scale rows, packed weight gradients and the activation gradients that are accumulate-only
        self.bwd_arena[:max(self.bwd_zero_bytes, 1)].zero_()
        if self.side_wgrad and self.side_stream is None:
            self.side_stream = torch.cuda.Stream(device=self.device)
        grads = {}
        # head gradients are GEMM operands (fp16): scale them by a power of two derived on the device from max|dp|
        live = [dp for dp in dps if dp is not None]
        amax = torch.stack([dp.detach().abs().max() for dp in live]).max().clamp(min=1e-30)
        # 2^floor(log2(4096 / amax)) = the value with its mantissa bits cleared (no exp2 / log2: torch compiles those
        # through NVRTC on first use, which every rank of a multi-GPU job would do at the same moment)
        hs = ((4096.0 / amax).float().view(torch.int32) & 0x7F800000).view(torch.float32).clamp(1e-30, 1e30)
        self.head_scale = torch.stack([hs, 1.0 / hs]).contiguous()           # device [s, 1/s]
        for (m, raw), dp in zip(self.yolo, dps):
            g = self.grad_of[id(raw)]
            if dp is None:
                continue
            call("b2y_yolo_grad_to_raw", ptr(dp.contiguous().float()), ptr(g.buf), HEAD_PAD, self.B, m.na, m.no, raw.H,
                 raw.W, S, ptr(self.head_scale), ops._gdt(g.buf), stream_ptr())
        G = lambda t: self.grad_of[id(t)]
        for idx in range(len(self.order) - 1, -1, -1):
            st = self.order[idx]
            kind = st[0]
            if kind == 'conv':
                self._conv_backward(st[1], grads, S, inv)
            elif kind == 'se':
                e = st[1]
                ops.se_bwd(e['src'].view(), G(e['y']).view(), e['fc1'].weight.detach(), e['fc2'].weight.detach(),
                           e['ws'], G(e['src']).view(), accumulate=True, dw1=self._grad_dst(e['fc1'].weight),
                           dw2=self._grad_dst(e['fc2'].weight), grad_scale=inv, ws_bwd=e['ws_bwd'])
            elif kind == 'add':
                _, first, others, out = st
                go = G(out).view()
                for j, s in enumerate([first] + list(others)):
                    gs = G(s).view()
                    n = min(gs.shape[3], go.shape[3])      # sliced addends receive the matching gradient channels
                    if self.first_write.get(('add', idx, j)):
                        ops.copy_channels(go[..., :n], gs[..., :n])
                    else:
                        ops.add(gs[..., :n], go[..., :n], out=gs[..., :n])
            elif kind == 'copy':
                _, srct, dst, off = st
                gd = G(dst)
                gs = G(srct).view()
                if self.first_write.get(('copy', idx)):
                    ops.copy_channels(gd.buf[..., gd.c0 + off:gd.c0 + off + srct.C], gs)
                else:
                    ops.add(gs, gd.buf[..., gd.c0 + off:gd.c0 + off + srct.C], out=gs)
            elif kind == 'upsample':
                _, src, out, s = st
                gy, gx = G(out).view(), G(src).view()
                call("b2y_upsample_nearest_bwd", ptr(gy), ops._pitch(gy), ptr(gx), ops._pitch(gx), self.B, src.H,
                     src.W, src.C, int(s), ops._gdt(gy), stream_ptr())
            elif kind == 'maxpool':
                _, src, out, k, s, tiny = st
                gy, gx, xv = G(out).view(), G(src).view(), src.view()
                call("b2y_maxpool_bwd", ptr(xv), ops._pitch(xv), ptr(gy), ops._pitch(gy), ptr(gx), ops._pitch(gx),
                     self.B, src.H, src.W, src.C, int(k), int(s), 1 if tiny else 0, ops._gdt(gy), stream_ptr())
        if self.side_wgrad:
            torch.cuda.current_stream().wait_stream(self.side_stream)      # join: every weight gradient is complete
        self._unpack_all()
        if self.own_grads:
            # plain-autograd path: hand out copies (the plan-owned buffers are overwritten by the next step)
            by_id = {id(p): p for p in self.params}
            for pid, g in self.own_grads.items():
                grads[by_id[pid]] = g
        return grads

    def _conv_backward(self, r, grads, S, inv):
        B = self.B
        gy = self.grad_of[id(r.y)]
        conv, bn = r.conv, r.bn
        Ho, Wo = r.y.H, r.y.W
        if bn is not None:
            dy = gy.view()
            if r.res is not None:   # fused shortcut: the same gradient also flows to the skip source
                gs = self.grad_of[id(r.res)].view()
                if self.first_write.get(('res', r.i)):
                    ops.copy_channels(dy, gs)
                else:
                    ops.add(gs, dy, out=gs)
            if self.side_wgrad:
                dz = self.dz_bufs.get(r.i)
                if dz is None:
                    # pitch = Cout rounded up to the K granule of the data-gradient GEMM; the pad channels stay zero
                    full = torch.zeros((B, Ho, Wo, r.Cpad), dtype=torch.float16, device=self.device)
                    dz = self.dz_bufs[r.i] = full[..., :r.Cout]
            else:
                if r.Cpad != r.Cout:
                    raise NotImplementedError("B2Y_WGRAD_STREAM=0 needs channel counts that are multiples of 16")
                dz = self.dz_scratch[:B * Ho * Wo * r.Cout].view(B, Ho, Wo, r.Cout)
            aux = self.dz_aux[r.aux_row]
            inv_s = aux[2:3]                                   # device scalar 1/s of this layer's dz
            if r.fast_bn:
                z = r.z.view()
                pixels = B * Ho * Wo
                a = ACT[r.act]
                call("b2y_bn_train_bwd_reduce", ptr(z), ops._pitch(z), ptr(dy), ops._pitch(dy), ptr(r.save),
                     ptr(r.sums), ptr(aux), pixels, r.Cout, a, float(r.slope), ops._gdt(dy), stream_ptr())
                call("b2y_bn_train_bwd_apply", ptr(z), ops._pitch(z), ptr(dy), ops._pitch(dy),
                     ptr(bn.weight.detach()), ptr(r.save), ptr(r.sums), ptr(dz), ops._pitch(dz), pixels, r.Cout, a,
                     float(r.slope), ops._gdt(dy), ptr(aux), C.c_void_p(aux.data_ptr() + 4),
                     ptr(self._grad_dst(bn.weight)), ptr(self._grad_dst(bn.bias)), float(inv), stream_ptr())
            else:
                dgb = self.dgb_scratch[:, :r.Cout]
                dgb.zero_()
                ops.bn_act_bwd(r.z.view(), dy, r.scale, r.shift, bn.weight.detach(), r.mean, r.invstd, r.act, r.slope,
                               dx=dz, dgamma=dgb[0], dbeta=dgb[1], aux=aux)
                self._emit(grads, bn.weight, dgb[0], inv)
                self._emit(grads, bn.bias, dgb[1], inv)
        else:
            dz = gy.buf[..., :r.Cpad] if r.head else gy.view()
            inv_s = self.head_scale[1:2]
            if r.act != 'linear':
                raise NotImplementedError("activation without BatchNorm in training")
            if conv.bias is not None:
                db = self.dgb_scratch[1, :r.Cpad]
                db.zero_()
                ops.bias_act_bwd_reduce(dz, dz, r.ones, r.zeros, 'linear', dbeta=db)
                self._emit(grads, conv.bias, db[:r.Cout] * inv_s, inv)
        I = conv.in_channels
        if r.stem and getattr(r, 'wstem', None) is not None and getattr(self, 'stem_ws', None) is not None:
            # dW = dz^T . im2col(x): the pixel-dimension GEMM on the tensor cores over the forward's workspace
            gw = self._grad_dst(conv.weight)
            kk = I * r.k * r.k
            xcol = self.stem_ws[:B * Ho * Wo * 32].view(B, Ho, Wo, 32)
            dwp = self.stem_dw[:r.Cpad * 32].view(r.Cpad, 1, 1, 32)
            dwp.zero_()
            ops.conv2d_bwd_weight(xcol, dz, 1, 1, 0, scale=inv, dw=dwp, inv_scale=inv_s)
            gw.copy_(dwp[:r.Cout, 0, 0, :kk].view(r.Cout, r.k, r.k, I).permute(0, 3, 1, 2))
        elif r.stem:
            gw = self._grad_dst(conv.weight)
            gw.zero_()
            d = ConvDesc(B, self.H, self.W, I, I, r.Cout, r.k, r.s, r.p, Ho, Wo, ops._pitch(dz), 0, 0.0, OUT_F16, 0)
            call("b2y_stem_conv_bwd_weight", C.byref(d), ptr(self.x), ptr(dz), ptr(gw), inv, ops._gdt(dz), stream_ptr())
            gw.mul_(inv_s)
        elif r.depthwise:
            xs = r.src.view()
            ops.dwconv2d_bwd_weight(xs, dz, r.k, r.s, r.p, alpha=inv, dw=r.dwg, inv_scale=inv_s)
            ops.axpby(r.dwg.reshape(-1), self._grad_dst(conv.weight).reshape(-1), 1.0, 0.0)
            gx = self.grad_of[id(r.src)]
            ops.dwconv2d_bwd_data(dz, conv.weight.detach(), (B, r.src.H, r.src.W, I), r.k, r.s, r.p, out=gx.view(),
                                  accumulate=True, inv_scale=inv_s)
        else:
            # packed fp32 gradient slice (zeroed with the backward arena); unpacked for all layers at the end.
            # K operands are taken with their zero pad: x with Ipad channels, dz with Cpad channels.
            xk = r.src.view_k()
            dzk = dz if dz.shape[3] == r.Cpad else _widen(dz, r.Cpad)
            if self.side_wgrad:
                main = torch.cuda.current_stream()
                self.side_stream.wait_stream(main)              # dz (and 1/s) of this layer are complete
                with torch.cuda.stream(self.side_stream):
                    ops.conv2d_bwd_weight(xk, dzk, r.k, r.s, r.p, scale=inv, dw=r.dw, inv_scale=inv_s)
            else:
                ops.conv2d_bwd_weight(xk, dzk, r.k, r.s, r.p, scale=inv, dw=r.dw, inv_scale=inv_s)
            gx = self.grad_of[id(r.src)]
            ops.conv2d_bwd_data(dzk, r.wT, (B, r.src.H, r.src.W, I), r.k, r.s, r.p, out=gx.view(),
                                accumulate=not self.first_write.get(('dgrad', r.i), False), inv_scale=inv_s)

    def _emit(self, grads, param, src, alpha):
        ops.axpby(src.contiguous(), self._grad_dst(param), alpha, 0.0)

    # ---------------------------------------------------------------------------------------------------------
    def run(self, x):
        """autograd entry: returns (yolo_out list, feature_out)."""
        outs = _DarknetTrain.apply(self, x, *self.params)
        # the running statistics were just updated through raw pointers (also under graph replay): eval plans re-fold
        object.__setattr__(self.model, '_b2y_epoch', getattr(self.model, '_b2y_epoch', 0) + 1)
        feats = LazyFeatures([None if t is None else t.view() for t in self.feature_views])
        return list(outs), feats


class _DarknetTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, *params):
        ctx.plan = plan
        ctx.params = params
        with torch.no_grad():
            outs = plan.forward_graphed(x)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dps):
        plan = ctx.plan
        with torch.no_grad():
            grads = plan.backward_graphed(dps)
        out = []
        for p in ctx.params:
            g = grads.get(p)
            out.append(g if (g is not None and p.requires_grad) else None)
        return (None, None) + tuple(out)
