"""Sync-free execution of a whole sparse backbone (eval mode).

The module API (spconv/conv.py) runs the way the reference does: one rulebook build with a host
round-trip per strided conv, conv + bias, then BatchNorm1d and ReLU as separate kernels
(pcdet/ops/spconv/conv.py:113-229, modules.py:125-137).  The engine runs the same layer graph B200-style:

  * geometry pre-pass -- rulebooks depend on coordinates only, so all 8/9 of them are built first, level by
    level, with every row count kept in a device scalar (SURVEY.md section 7 "data-dependent shapes");
  * feature pass -- one fused kernel per conv layer: gather + contraction + bias + folded BatchNorm
    (+ residual) + ReLU, reading the output-major neighbour map;
  * ONE device->host copy of the five row counts at the end, to hand out correctly shaped tensors.

Nothing in between touches the host, so the whole step can be captured in a CUDA graph (pipeline.py).
All buffers live in a grow-only arena sized by capacity bounds; the library itself never allocates.
"""
import numpy as np
import torch
from torch import nn

from . import _lib
from .spconv import ops
from .spconv.conv import SparseConvolution
from .spconv.modules import SparseSequential
from .spconv.structure import SparseConvTensor


class _Step(object):
    __slots__ = ("conv", "bn", "relu", "key", "subm", "in_level", "out_level", "in_buf", "out_buf", "res_buf",
                 "export", "cin", "cout", "kvol")


class _Book(object):
    __slots__ = ("key", "subm", "in_level", "out_level", "ksize", "stride", "pad", "dil", "kvol", "fanout")


def fold_bn(bn):
    """Eval BatchNorm1d as y = x*scale + shift (fp32): scale = gamma/sqrt(var+eps), shift = beta - mean*scale."""
    with torch.no_grad():
        inv = torch.rsqrt(bn.running_var.float() + bn.eps)
        g = bn.weight.float() if bn.weight is not None else torch.ones_like(inv)
        b = bn.bias.float() if bn.bias is not None else torch.zeros_like(inv)
        scale = (g * inv).contiguous()
        shift = (b - bn.running_mean.float() * scale).contiguous()
    return scale, shift


def trace_backbone(net):
    """Flattens a VoxelBackBone8x-style module tree into fused conv steps.

    Recognised patterns: SparseSequential(SparseConvolution, BatchNorm1d, ReLU) (post_act_block,
    spconv_backbone.py:10-29) and SparseBasicBlock (spconv_backbone.py:32-68).  Returns None when the tree
    contains anything else, in which case callers use the plain module path.
    """
    from .spconv_backbone import SparseBasicBlock
    steps, books, levels = [], {}, [list(int(s) for s in net.sparse_shape)]
    exports = {"conv1": "x_conv1", "conv2": "x_conv2", "conv3": "x_conv3", "conv4": "x_conv4", "conv_out": "out"}
    state = dict(level=0, buf=-1, nbuf=0)  # buf -1 = the caller's voxel_features

    def add_conv(conv, bn, relu, res_buf):
        if conv.transposed or conv.inverse or conv.conv1x1 or conv.ndim != 3:
            return False
        key = conv.indice_key if conv.indice_key is not None else "_anon%d" % len(steps)
        if key in books:
            bk = books[key]
            if bk.in_level != state["level"]:
                return False
        else:
            bk = _Book()
            bk.key, bk.subm, bk.in_level = key, conv.subm, state["level"]
            bk.ksize = list(conv.kernel_size)
            bk.dil = list(conv.dilation)
            bk.kvol = int(np.prod(bk.ksize))
            if conv.subm:
                bk.stride, bk.pad = [1, 1, 1], [k // 2 for k in bk.ksize]  # spconv_ops.h:76-80
                bk.out_level, bk.fanout = state["level"], 1
            else:
                bk.stride, bk.pad = list(conv.stride), list(conv.padding)
                out_shape = ops.get_conv_output_size(levels[state["level"]], bk.ksize, bk.stride, bk.pad, bk.dil)
                levels.append([int(s) for s in out_shape])
                bk.out_level = len(levels) - 1
                bk.fanout = ops.candidate_fanout(bk.ksize, bk.stride, bk.pad, bk.dil)
            if bk.kvol > _lib.MAX_KVOL:
                return False
            books[key] = bk
        st = _Step()
        st.conv, st.bn, st.relu, st.key, st.subm = conv, bn, relu, key, conv.subm
        st.in_level, st.out_level = bk.in_level, bk.out_level
        st.in_buf, st.res_buf = state["buf"], res_buf
        st.out_buf = state["nbuf"]
        st.cin, st.cout, st.kvol = conv.in_channels, conv.out_channels, bk.kvol
        st.export = None
        state["nbuf"] += 1
        state["buf"], state["level"] = st.out_buf, bk.out_level
        steps.append(st)
        return True

    def walk(seq):
        mods = list(seq._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, SparseBasicBlock):
                if m.downsample is not None:
                    return False
                identity = state["buf"]
                if identity < 0:
                    return False
                if not add_conv(m.conv1, m.bn1, True, None):
                    return False
                if not add_conv(m.conv2, m.bn2, True, identity):
                    return False
                i += 1
            elif isinstance(m, SparseConvolution):
                bn = relu = None
                j = i + 1
                if j < len(mods) and isinstance(mods[j], nn.BatchNorm1d):
                    bn = mods[j]
                    j += 1
                if j < len(mods) and isinstance(mods[j], nn.ReLU):
                    relu = True
                    j += 1
                if not add_conv(m, bn, bool(relu), None):
                    return False
                i = j
            elif isinstance(m, SparseSequential):
                if not walk(m):
                    return False
                i += 1
            else:
                return False
        return True

    for name, child in net.named_children():
        if not isinstance(child, SparseSequential) or not walk(child):
            return None
        if name in exports and steps:
            steps[-1].export = exports[name]
    return steps, list(books.values()), levels


class _LazyIndiceDict(dict):
    """indice_dict whose values are produced on first access (the reference-layout pair tensors are not needed by
    anything on the hot path; consumers that read them - `find_indice_pair`, `indice_dict[key]` - get them built then)."""

    def defer(self, key, thunk):
        dict.__setitem__(self, key, _Deferred(thunk))

    def _resolve(self, key, value):
        if isinstance(value, _Deferred):
            value = value.thunk()
            dict.__setitem__(self, key, value)
        return value

    def __getitem__(self, key):
        return self._resolve(key, dict.__getitem__(self, key))

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


class _Deferred(object):
    __slots__ = ("thunk",)

    def __init__(self, thunk):
        self.thunk = thunk


class CapacityOverflow(RuntimeError):
    """A level produced more rows than its arena buffers hold (nothing was written out of bounds).  Raised when the
    row counts reach the host; `BackboneEngine.grow()` enlarges the bounds, then the step is run again."""


class BackboneEngine(object):
    """Runs a traced backbone.  precision: 'fp32' (fp32 storage, fp32-accurate arithmetic) or 'bf16'
    (bf16 storage, fp32 accumulation; the entry layer reads fp32 voxel features)."""

    SIDE_STREAMS = 4

    def __init__(self, net, precision="fp32", materialize_pairs=True, use_tensor_cores=True, sort_rows=True,
                 concurrent=True, cap_growth='auto'):
        traced = trace_backbone(net)
        if traced is None:
            raise NotImplementedError("backbone layout not recognised by the fused engine")
        self.steps, self.books, self.level_shapes = traced
        self.net = net
        self.precision = precision
        # True: the reference-layout pair tensors are built with every step (on a side stream); 'lazy': when
        # indice_dict[key] is first read; False: never (indice_dict stays empty)
        self.materialize_pairs = materialize_pairs
        self.use_tensor_cores = use_tensor_cores
        # grouped row order for the tensor-core layers (same results, fewer stages): True = every rulebook, 'subm' =
        # the submanifold ones only (a strided rulebook feeds ONE layer: grouping it costs more than it saves), False
        self.sort_rows = sort_rows
        self.concurrent = concurrent  # geometry and feature pass on forked streams (joined before launch returns)
        # Row capacity of level l+1 = min(cap_growth * capacity of level l, hard bound).  The hard bound
        # (fan-out 8 per strided conv, or the dense volume) is 10-100x what LiDAR frames produce - a KITTI batch
        # of 8 would reserve ~28 GB of rulebooks and feature buffers against ~2 GB at growth 2 - so the arena
        # starts from the modest bound and grows on CapacityOverflow.  None = hard bound from the start.
        self.cap_growth = cap_growth
        self._side = None
        self.arena = None
        self.arena_gen = 0
        self._param_key = None
        self._params = None
        self._tensors = None
        # liveness of feature buffers: last step reading each one (exports live forever)
        last_read = {}
        for i, st in enumerate(self.steps):
            last_read[st.in_buf] = i
            if st.res_buf is not None:
                last_read[st.res_buf] = i
        self._last_read = last_read

    # ------------------------------------------------------------------ parameters
    def param_key(self, device):
        """Cheap identity of everything the packed weights / folded BatchNorm were derived from: (address, version
        counter) of every parameter and buffer.  In-place updates (load_state_dict, optimizer steps, .copy_) bump
        the version; replaced tensors change the address.  Recomputed on every launch AND every graph replay."""
        if self._tensors is None:  # walking the module tree costs ~0.4 ms, the tensors themselves ~30 us
            self._tensors = list(self.net.parameters()) + list(self.net.buffers())
        return tuple([t.data_ptr() for t in self._tensors] + [t._version for t in self._tensors]) + \
            (str(device), self.precision)

    def _prepare_params(self, device, rescan=False):
        """rescan: walk the module tree again (Parameter objects replaced, not just updated in place)."""
        if rescan:
            self._tensors = None
        key = self.param_key(device)
        if key == self._param_key:
            return self._params
        prm = []
        lib = _lib.load()
        for st in self.steps:
            w = st.conv.weight.detach()
            if not w.is_cuda:
                raise ValueError("backbone parameters must live on the CUDA device")
            w = w.float().reshape(st.kvol, st.cin, st.cout).contiguous()
            bias = st.conv.bias.detach().float().contiguous() if st.conv.bias is not None else None
            scale = shift = None
            if st.bn is not None:
                scale, shift = fold_bn(st.bn)
                if bias is not None:
                    # (acc + bias) * scale + shift == acc * scale + (bias * scale + shift): one fused multiply-add and
                    # two parameter loads per output element in the epilogue instead of three
                    shift = (bias * scale + shift).contiguous()
                    bias = None
            mode = self._mode_for(st)
            packed = None
            if mode in (_lib.MODE_BF16_TC, _lib.MODE_TF32X3_TC):
                nbytes = lib.fv2p_pack_weight_bytes(st.kvol, st.cin, st.cout, mode)
                packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
                _lib.check(lib.fv2p_pack_weight(_lib.ptr(w), st.kvol, st.cin, st.cout, mode, _lib.ptr(packed),
                                                _lib.stream_ptr(device)), "pack_weight")
            prm.append(dict(w=w, packed=packed, bias=bias, scale=scale, shift=shift, mode=mode))
        self._params, self._param_key = prm, key
        return prm

    def _mode_for(self, st):
        first = st.in_buf < 0
        tc_ok = self.use_tensor_cores and st.cin % 16 == 0 and st.cout % 16 == 0 and st.cin <= 128 and \
            16 <= st.cout <= 128 and _tc_available()
        if self.precision == "bf16":
            if first:
                return _lib.MODE_F32_IN_BF16_OUT
            return _lib.MODE_BF16_TC if tc_ok else _lib.MODE_BF16_SIMT
        return _lib.MODE_TF32X3_TC if (tc_ok and not first) else _lib.MODE_F32

    # ------------------------------------------------------------------ arena
    # Row capacity of a level relative to the level it is built from, when cap_growth == 'auto': what LiDAR-like
    # frames produce (x1.06-1.13 at the first stride-2 conv, then x0.45-0.57, x0.40, x0.85) with 30-100 %
    # head room.  A step that overflows sets a status bit, the bounds grow and the step runs again.
    AUTO_GROWTH_FIRST, AUTO_GROWTH_DEEPER, AUTO_GROWTH_NARROW = 1.5, 0.6, 1.0

    def _growth(self, bk):
        if not self.cap_growth:
            return float(bk.fanout)
        if self.cap_growth == 'auto':
            if bk.fanout <= 2:
                return min(float(bk.fanout), self.AUTO_GROWTH_NARROW)
            return self.AUTO_GROWTH_FIRST if bk.in_level == 0 else min(float(bk.fanout), self.AUTO_GROWTH_DEEPER)
        return min(float(bk.fanout), float(self.cap_growth))

    def _caps(self, cap0, batch):
        caps = [int(cap0)]
        for bk in self.books:
            if not bk.subm:
                vol = int(np.prod(self.level_shapes[bk.out_level])) * int(batch)
                caps.append(max(1, min(int(caps[bk.in_level] * self._growth(bk)) + 1024,
                                       caps[bk.in_level] * bk.fanout, vol)))
        return caps

    def grow(self):
        """Enlarges the capacity bounds after a CapacityOverflow ('auto' -> x2 per level -> x8 -> the hard bound); the
        next launch allocates a new arena.  Returns False when the bounds are already the hard ones."""
        if not self.cap_growth:
            return False
        if self.cap_growth == 'auto':
            self.cap_growth = 2.0
        else:
            self.cap_growth = None if self.cap_growth >= 8 else self.cap_growth * 4.0
        self.arena = None
        self.arena_gen += 1  # captured graphs of the old arena are stale
        return True

    @staticmethod
    def _geo_key(bk):
        return (bool(bk.subm), bk.in_level, tuple(bk.ksize), tuple(bk.stride), tuple(bk.pad), tuple(bk.dil))

    def _ensure_arena(self, device, cap0, batch):
        a = self.arena
        if a is not None and a["device"] == device and a["batch"] >= batch and a["caps"][0] >= cap0:
            return a
        cap0 = int(cap0 * 1.15) + 256 if a is not None else int(cap0)
        caps = self._caps(cap0, batch)
        lib = _lib.load()
        dt = torch.float32 if self.precision == "fp32" else torch.bfloat16
        a = dict(device=device, batch=batch, caps=caps)
        a["counts"] = torch.zeros((len(caps) + 1,), dtype=torch.int32, device=device)  # [levels..., status]
        a["counts_host"] = torch.zeros((len(caps) + 1,), dtype=torch.int32).pin_memory()
        a["indices"] = [None] + [torch.empty((c, 4), dtype=torch.int32, device=device) for c in caps[1:]]
        # one coordinate table per level: level 0 is built from the voxel coordinates, level l+1 is the output table of
        # the strided rulebook l -> l+1 and then serves the submanifold rulebooks of level l+1
        a["tables"] = [torch.empty(lib.fv2p_table_bytes(c) + 16, dtype=torch.uint8, device=device) for c in caps]
        tc_modes = (_lib.MODE_BF16_TC, _lib.MODE_TF32X3_TC)
        tc_books = {st.key for st in self.steps if self._mode_for(st) in tc_modes}
        books, geo = {}, {}
        for bk in self.books:
            gk = self._geo_key(bk)
            d = geo.get(gk)
            if d is None:
                cin_cap, cout_cap = caps[bk.in_level], caps[bk.out_level]
                # neighbour-map rows padded to whole 128-row tiles: the conv's bulk copies need 16-byte aligned rows
                cols = (cout_cap + 127) // 128 * 128
                d = dict(geo=gk, book=bk, keys=[], nbr=torch.empty((bk.kvol, cols), dtype=torch.int32, device=device),
                         sorted=False, pairs=None, pair_num=None, pair_ws=None)
                if not bk.subm:
                    d["ws"] = torch.empty(lib.fv2p_conv_neighbours_workspace_bytes(cin_cap, bk.kvol) + 256,
                                          dtype=torch.uint8, device=device)
                geo[gk] = d
            d["keys"].append(bk.key)
            if self._groups(bk) and bk.key in tc_books and not d["sorted"]:
                cout_cap = caps[bk.out_level]
                cols = d["nbr"].shape[1]
                d["sorted"] = True
                d["perm"] = torch.empty((cout_cap,), dtype=torch.int32, device=device)
                d["nbr_sorted"] = torch.empty((bk.kvol, cols), dtype=torch.int32, device=device)
                d["tile_order"] = torch.empty((cout_cap // 128 + 2, 2), dtype=torch.int32, device=device)
                d["group_ws"] = torch.empty(lib.fv2p_group_rows_workspace_bytes(cout_cap) + 256, dtype=torch.uint8,
                                            device=device)
            books[bk.key] = d
        a["books"], a["geo"] = books, list(geo.values())
        # everything the geometry pass wants cleared / zeroed / set to -1 before it starts: ONE launch per step
        items = [(_lib.PREFILL_TABLE, t, c, 0) for t, c in zip(a["tables"], caps)]
        for d in a["geo"]:
            bk = d["book"]
            if bk.subm:
                if all(k % 2 == 1 for k in bk.ksize) and all(x == 1 for x in bk.dil) and bk.kvol > 1:
                    items.append((_lib.PREFILL_NBR_MIRROR, d["nbr"], bk.kvol, d["nbr"].shape[1]))
            else:
                items.append((_lib.PREFILL_NBR_ALL, d["nbr"], bk.kvol, d["nbr"].shape[1]))
                items.append((_lib.PREFILL_CONV_WS, d["ws"], caps[bk.in_level], bk.kvol))
            if d["sorted"]:
                items.append((_lib.PREFILL_GROUP_WS, d["group_ws"], caps[bk.out_level], 0))
        a["prefill"] = _lib.prefill_items(items)
        a["prefill_no_table0"] = _lib.prefill_items(items[1:])  # the voxelizer clears and builds the level-0 table
        # two scheduler words per conv step (tile counter, CTAs done); zero between launches, the kernel re-arms them
        a["sched"] = torch.zeros((len(self.steps), 2), dtype=torch.int32, device=device)
        # feature buffers with liveness-based reuse
        bufs, free, owner = {}, [], {}
        for i, st in enumerate(self.steps):
            shape = (caps[st.out_level], st.cout)
            pick = None
            for j, (t, last, exported) in enumerate(free):
                if not exported and last < i and tuple(t.shape) == shape and st.in_buf != owner[j] and \
                        st.res_buf != owner[j]:
                    pick = j
                    break
            if pick is None:
                t = torch.empty(shape, dtype=dt, device=device)
                free.append([t, 0, False])
                pick = len(free) - 1
            free[pick][1] = self._last_read.get(st.out_buf, i)
            free[pick][2] = st.export is not None
            owner[pick] = st.out_buf
            bufs[st.out_buf] = free[pick][0]
        a["bufs"] = bufs
        self.arena = a
        self.arena_gen += 1  # captured CUDA graphs hold arena addresses: a new arena invalidates them
        return a

    def arena_bytes(self):
        """Device bytes held by the current arena (rulebooks, tables, feature buffers)."""
        a = self.arena
        if a is None:
            return 0
        seen, total = set(), 0

        def add(t):
            nonlocal total
            if isinstance(t, torch.Tensor) and t.is_cuda and t.data_ptr() not in seen:
                seen.add(t.data_ptr())
                total += t.numel() * t.element_size()

        for t in a["indices"] + a["tables"] + list(a["bufs"].values()) + [a["counts"], a["sched"]]:
            add(t)
        for d in a["geo"]:
            for v in d.values():
                add(v)
        return total

    # ------------------------------------------------------------------ run
    def _streams(self, device):
        if self._side is None or self._side[0].device != device:
            self._side = [torch.cuda.Stream(device=device) for _ in range(self.SIDE_STREAMS + 3)]
        return self._side[:-3], self._side[-3], self._side[-2], self._side[-1]

    def prefill(self, device, cap0, batch_size, table0_external=False):
        """Enqueues the step's ONE clearing launch (tables, scan states, -1 fills) on a side stream forked from the
        current stream, so that it runs next to whatever the caller enqueues next (the voxelizer).  Returns the event
        to hand to launch(prefilled=...).  Nothing of the previous step may still be reading the arena on another
        stream (launch() joins its streams before it returns)."""
        a = self._ensure_arena(device, max(int(cap0), 1), int(batch_size))
        if not self.concurrent:
            return None
        s_fill = self._streams(device)[3]
        main = torch.cuda.current_stream(device)
        ev = torch.cuda.Event()
        ev.record(main)
        s_fill.wait_event(ev)
        items, n_items = a["prefill_no_table0" if table0_external else "prefill"]
        with torch.cuda.device(device), torch.cuda.stream(s_fill):
            _lib.check(_lib.load().fv2p_geometry_prefill(items, n_items, _lib.stream_ptr(device)), "geometry_prefill")
        done = torch.cuda.Event()
        done.record(s_fill)
        return done

    def level0_table(self, device, cap0, batch_size):
        """(table buffer, row capacity, spatial shape) of the level-0 coordinate table, for a voxelizer that builds it
        while it assigns the voxel rows (BatchVoxelizer(level0_table=...), then launch(table0_built=True))."""
        a = self._ensure_arena(device, max(int(cap0), 1), int(batch_size))
        return a["tables"][0], a["caps"][0], [int(v) for v in self.level_shapes[0]]

    def launch(self, voxel_features, voxel_coords, batch_size, n0_dev=None, cap0=None, features_ready=None,
               prefilled=None, run_convs=True, table0_built=False):
        """Enqueues geometry + feature passes on the current stream.  No host sync.
        ``features_ready``: event after which voxel_features may be read (None: already ordered on this stream).
        ``prefilled``: event from prefill() when the clearing launch was already enqueued (None: done here).
        ``run_convs``: False enqueues the geometry pass only (bench.py times it on its own that way).
        ``table0_built``: the level-0 coordinate table already holds these coordinates (see level0_table()).

        voxel_features [>=cap0, F] fp32, voxel_coords [>=cap0, 4] int32, both CUDA and contiguous;
        live row count = *n0_dev (device int32) if given, else cap0 (defaults to voxel_coords.shape[0]).
        """
        dev = _lib.require_device(voxel_features)
        device = voxel_features.device
        if voxel_coords.dtype != torch.int32 or voxel_features.dtype != torch.float32:
            raise ValueError("engine expects float32 features and int32 coordinates")
        if not (voxel_features.is_contiguous() and voxel_coords.is_contiguous()):
            raise ValueError("engine expects contiguous inputs")
        cap0 = int(voxel_coords.shape[0] if cap0 is None else cap0)
        a = self._ensure_arena(device, max(cap0, 1), int(batch_size))
        prm = self._prepare_params(device, rescan=not torch.cuda.is_current_stream_capturing())
        lib = _lib.load()
        counts = a["counts"]
        caps = a["caps"]
        PRE = _lib.FLAG_PREFILLED
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(device)
            if self.concurrent:
                side, s_pairs, s_conv, _ = self._streams(device)
            else:
                side, s_pairs, s_conv = [main], main, main
            counts[len(caps):].zero_()
            if n0_dev is None:
                counts[0:1].fill_(cap0)
                n0_dev = counts[0:1]
            else:
                counts[0:1].copy_(n0_dev.view(-1)[0:1])
            level_ind = [voxel_coords] + a["indices"][1:]
            # Row capacities as the ARENA knows them, for every geometry call: the library carves its workspaces by the
            # capacity it is given, and the prefill items were computed from these.  (The voxel buffers may hold fewer
            # rows than caps[0]; the kernels only touch live rows, and the live count never exceeds either bound.)
            level_cap = list(caps)
            n_ptr = [_lib.ctypes.c_void_p(counts.data_ptr() + 4 * i) for i in range(len(caps))]
            status_ptr = _lib.ctypes.c_void_p(counts.data_ptr() + 4 * len(caps))

            # Dependency chains, forked from and joined back into the caller's stream (so the whole thing is still one
            # stream-ordered step, and one CUDA graph when captured):
            #   main   ONE prefill launch, the level-0 table, then the strided rulebooks level by level (each needs
            #          the previous level's output coordinates)
            #   side   the submanifold rulebooks and every row grouping: leaves of that chain, independent of each
            #          other, dealt round-robin to SIDE_STREAMS streams
            #   s_pairs the reference-layout pair lists when they are materialised eagerly (nothing in the step waits
            #          for them but the final join)
            #   s_conv the feature pass, each layer waiting only for what it reads (the grouped set for the
            #          tensor-core kernels, the plain neighbour map otherwise)
            def mark(stream):
                ev = torch.cuda.Event()
                ev.record(stream)
                return ev

            if prefilled is not None:
                main.wait_event(prefilled)
            else:
                items, n_items = a["prefill_no_table0" if table0_built else "prefill"]
                _lib.check(lib.fv2p_geometry_prefill(items, n_items, _lib.stream_ptr(device)), "geometry_prefill")
            if not table0_built:
                _lib.check(lib.fv2p_table_build(_lib.ptr(voxel_coords), cap0, n_ptr[0],
                                                _lib.i32x3(self.level_shapes[0]), _lib.ptr(a["tables"][0]), caps[0],
                                                status_ptr, PRE, _lib.stream_ptr(device)), "table_build")
            level_ready = {0: mark(main)}
            built, grouped = {}, {}
            forked = []  # streams that joined this step (a captured step may only be joined by streams it forked)
            for j, d in enumerate(a["geo"]):
                bk = d["book"]
                s_side = side[j % len(side)]
                if s_side not in forked:
                    forked.append(s_side)
                if bk.subm:
                    s_side.wait_event(level_ready[bk.in_level])
                    with torch.cuda.stream(s_side):
                        st = lib.fv2p_subm_neighbours(_lib.ptr(level_ind[bk.in_level]), level_cap[bk.in_level],
                                                      n_ptr[bk.in_level], int(batch_size),
                                                      _lib.i32x3(self.level_shapes[bk.in_level]), _lib.i32x3(bk.ksize),
                                                      _lib.i32x3(bk.dil), _lib.ptr(a["tables"][bk.in_level]),
                                                      caps[bk.in_level], _lib.ptr(d["nbr"]), d["nbr"].shape[1], PRE,
                                                      _lib.stream_ptr(device))
                    built[id(d)] = mark(s_side)
                else:
                    with torch.cuda.stream(main):
                        st = lib.fv2p_conv_neighbours(_lib.ptr(level_ind[bk.in_level]), level_cap[bk.in_level],
                                                      n_ptr[bk.in_level], int(batch_size),
                                                      _lib.i32x3(self.level_shapes[bk.out_level]),
                                                      _lib.i32x3(bk.ksize), _lib.i32x3(bk.stride), _lib.i32x3(bk.pad),
                                                      _lib.i32x3(bk.dil), _lib.ptr(level_ind[bk.out_level]),
                                                      level_cap[bk.out_level], n_ptr[bk.out_level],
                                                      _lib.ptr(a["tables"][bk.out_level]), caps[bk.out_level],
                                                      _lib.ptr(d["nbr"]), d["nbr"].shape[1], status_ptr,
                                                      _lib.ptr(d["ws"]), d["ws"].numel(), PRE, _lib.stream_ptr(device))
                    level_ready[bk.out_level] = built[id(d)] = mark(main)
                    s_side.wait_event(built[id(d)])
                _lib.check(st, "rulebook[%s]" % bk.key)
                if d["sorted"]:
                    with torch.cuda.stream(s_side):
                        st = lib.fv2p_group_rows(_lib.ptr(d["nbr"]), d["nbr"].shape[1], bk.kvol, int(bk.ksize[2]),
                                                 level_cap[bk.out_level], n_ptr[bk.out_level], _lib.ptr(d["perm"]),
                                                 _lib.ptr(d["nbr_sorted"]), d["nbr_sorted"].shape[1],
                                                 _lib.ptr(d["tile_order"]), _lib.ptr(d["group_ws"]),
                                                 d["group_ws"].numel(), PRE, _lib.stream_ptr(device))
                    _lib.check(st, "group_rows[%s]" % bk.key)
                    grouped[id(d)] = mark(s_side)
                if self.materialize_pairs is True:
                    if s_pairs not in forked:
                        forked.append(s_pairs)
                    s_pairs.wait_event(built[id(d)])
                    with torch.cuda.stream(s_pairs):
                        self._pairs_call(a, d, level_ind, level_cap, n_ptr, batch_size)
            waited = set()
            forked.append(s_conv)
            s_conv.wait_event(level_ready[0])
            if features_ready is not None:
                s_conv.wait_event(features_ready)
            for i, (st_, p) in enumerate(zip(self.steps, prm) if run_convs else ()):
                d = a["books"][st_.key]
                ev = grouped[id(d)] if self.conv_operands(a, st_, p)[1] is not None else built[id(d)]
                if id(ev) not in waited:
                    s_conv.wait_event(ev)
                    waited.add(id(ev))
                with torch.cuda.stream(s_conv):
                    self.run_conv_step(a, i, p, voxel_features, cap0)
            if self.concurrent:  # join: everything this step enqueued is ordered before what the caller does next
                for s_ in forked:
                    if s_ is not main:
                        main.wait_event(mark(s_))
        a["level_cap"] = level_cap
        a["step"] = a.get("step", 0) + 1
        return a

    def _pairs_call(self, a, d, level_ind, level_cap, n_ptr, batch_size):
        """Enqueues the reference-layout pair lists of one geometry on the current stream (buffers on first use)."""
        lib = _lib.load()
        bk = d["book"]
        device = a["device"]
        cin_cap = a["caps"][bk.in_level]
        if d["pairs"] is None:
            d["pairs"] = torch.empty((bk.kvol, 2, cin_cap), dtype=torch.int32, device=device)
            d["pair_num"] = torch.zeros((bk.kvol,), dtype=torch.int32, device=device)
            d["pair_ws"] = torch.empty(lib.fv2p_pairs_workspace_bytes(cin_cap, bk.kvol) + 256, dtype=torch.uint8,
                                       device=device)
        if bk.subm:
            st = lib.fv2p_subm_pairs(_lib.ptr(level_ind[bk.in_level]), level_cap[bk.in_level], n_ptr[bk.in_level],
                                     int(batch_size), _lib.i32x3(self.level_shapes[bk.in_level]), _lib.i32x3(bk.ksize),
                                     _lib.i32x3(bk.dil), _lib.ptr(a["tables"][bk.in_level]), a["caps"][bk.in_level],
                                     _lib.ptr(d["nbr"]), d["nbr"].shape[1], _lib.ptr(d["pairs"]), d["pairs"].shape[2],
                                     _lib.ptr(d["pair_num"]), _lib.ptr(d["pair_ws"]), d["pair_ws"].numel(),
                                     _lib.stream_ptr(device))
        else:
            st = lib.fv2p_conv_pairs(_lib.ptr(level_ind[bk.in_level]), level_cap[bk.in_level], n_ptr[bk.in_level],
                                     int(batch_size), _lib.i32x3(self.level_shapes[bk.out_level]),
                                     _lib.i32x3(bk.ksize), _lib.i32x3(bk.stride), _lib.i32x3(bk.pad),
                                     _lib.i32x3(bk.dil), _lib.ptr(a["tables"][bk.out_level]), a["caps"][bk.out_level],
                                     _lib.ptr(d["pairs"]), d["pairs"].shape[2], _lib.ptr(d["pair_num"]),
                                     _lib.ptr(d["pair_ws"]), d["pair_ws"].numel(), _lib.stream_ptr(device))
        _lib.check(st, "pairs[%s]" % bk.key)

    def _groups(self, bk):
        return bool(self.sort_rows) and (self.sort_rows != 'subm' or bk.subm)

    def conv_operands(self, a, step, prm):
        """(neighbour map, row order, tile order) a conv step reads: the mask-sorted set for the tensor-core modes."""
        d = a["books"][step.key]
        if d["sorted"] and prm["mode"] in (_lib.MODE_BF16_TC, _lib.MODE_TF32X3_TC):
            return d["nbr_sorted"], d["perm"], d["tile_order"]
        return d["nbr"], None, None

    def run_conv_step(self, a, i, prm, voxel_features, cap0):
        """Enqueues conv step ``i`` (fv2p_conv_fwd) on the current stream; its rulebook must already be enqueued."""
        st_ = self.steps[i]
        src = voxel_features if st_.in_buf < 0 else a["bufs"][st_.in_buf]
        res = a["bufs"][st_.res_buf] if st_.res_buf is not None else None
        out = a["bufs"][st_.out_buf]
        nbr, perm, order = self.conv_operands(a, st_, prm)
        w = prm["packed"] if prm["packed"] is not None else prm["w"]
        tc = prm["mode"] in (_lib.MODE_BF16_TC, _lib.MODE_TF32X3_TC)
        cap = cap0 if st_.out_level == 0 else a["caps"][st_.out_level]
        rc = _lib.load().fv2p_conv_fwd(
            _lib.ptr(src), src.shape[0], _lib.ptr(w), _lib.ptr(nbr), nbr.shape[1], _lib.ptr(perm), _lib.ptr(order),
            _lib.ptr(a["sched"][i]) if tc else None, st_.kvol, cap,
            _lib.ctypes.c_void_p(a["counts"].data_ptr() + 4 * st_.out_level), st_.cin, st_.cout,
            _lib.ptr(prm["bias"]), _lib.ptr(prm["scale"]), _lib.ptr(prm["shift"]), _lib.ptr(res), int(st_.relu),
            prm["mode"], _lib.ptr(out), _lib.stream_ptr(a["device"]))
        _lib.check(rc, "conv_fwd[%s]" % st_.key)

    def collect(self, a, voxel_coords, batch_size, sync=True):
        """One D2H copy of the row counts, then correctly shaped views (valid until the next launch)."""
        a["counts_host"].copy_(a["counts"], non_blocking=True)
        if sync:
            torch.cuda.current_stream(a["device"]).synchronize()
        return self.views(a, voxel_coords, batch_size)

    def views(self, a, voxel_coords, batch_size):
        host = a["counts_host"].tolist()
        n_levels = len(a["caps"])
        status = host[n_levels]
        if status:
            raise CapacityOverflow("fv2p_b200 engine: capacity overflow (status %d): a level has more rows than the "
                                   "arena bound (cap_growth=%s); call grow() and run the step again" %
                                   (status, self.cap_growth))
        n = host[:n_levels]
        level_ind = [voxel_coords[:n[0]]] + [t[:n[i + 1]] for i, t in enumerate(a["indices"][1:])]
        nbr_dict = {}
        indice_dict = _LazyIndiceDict() if self.materialize_pairs else {}
        for bk in self.books:
            d = a["books"][bk.key]
            nbr_dict[bk.key] = d["nbr"][:, :n[bk.out_level]]
            if self.materialize_pairs:
                indice_dict.defer(bk.key, self._pairs_thunk(a, d, bk, level_ind, n, voxel_coords, batch_size))
        outs = {}
        for st in self.steps:
            if st.export is None:
                continue
            t = SparseConvTensor(a["bufs"][st.out_buf][:n[st.out_level]], level_ind[st.out_level],
                                 self.level_shapes[st.out_level], batch_size)
            t.indice_dict, t.nbr_dict = indice_dict, nbr_dict
            # the level's coordinate table (rows of t.indices), for pointops.voxel_three_nn / voxel_query
            t.fv2p_table = (a["tables"][st.out_level], a["caps"][st.out_level])
            outs[st.export] = t
        return outs, n

    def _pairs_thunk(self, a, d, bk, level_ind, n, voxel_coords, batch_size):
        """indice_dict[key] = the reference's 5-tuple (conv.py:180-183).  With materialize_pairs=True the pair lists were
        enqueued with the step; with 'lazy' they are built from the neighbour map / the level's table when the entry is
        first read (valid, like every view, until the next launch)."""
        gen = self.arena_gen

        def make():
            if self.materialize_pairs is not True or d["pairs"] is None:
                if gen != self.arena_gen:
                    raise RuntimeError("indice_dict entry read after the arena it pointed into was replaced")
                if d.get("pairs_step") != a["step"]:  # two indice_keys over one geometry share the tensors
                    d["pairs_step"] = a["step"]
                    counts = a["counts"]
                    n_ptr = [_lib.ctypes.c_void_p(counts.data_ptr() + 4 * i) for i in range(len(a["caps"]))]
                    with torch.cuda.device(a["device"]):
                        self._pairs_call(a, d, [voxel_coords] + a["indices"][1:], a["level_cap"], n_ptr, batch_size)
            return (level_ind[bk.out_level], level_ind[bk.in_level], d["pairs"][:, :, :n[bk.in_level]], d["pair_num"],
                    self.level_shapes[bk.in_level])
        return make

    def launch_count(self, table0_built=False):
        """Kernels of libfv2p_b200 enqueued by one launch(): one prefill, the level-0 table (unless the voxelizer built
        it), 1 launch per submanifold
        and 3 per strided rulebook, 2 per row grouping, one fused conv per layer (+ the pair lists when they are
        materialised with the step: counts, scan, compaction, and the input-side probe for strided rulebooks)."""
        n = len(self.steps) + (1 if table0_built else 2)
        a = self.arena
        geo = a["geo"] if a is not None else None
        if geo is None:
            seen, geo = set(), []
            tc_modes = (_lib.MODE_BF16_TC, _lib.MODE_TF32X3_TC)
            tc_books = {st.key for st in self.steps if self._mode_for(st) in tc_modes}
            for bk in self.books:
                gk = self._geo_key(bk)
                if gk not in seen:
                    seen.add(gk)
                    geo.append(dict(book=bk, sorted=False))
                if self._groups(bk) and bk.key in tc_books:
                    [g for g in geo if self._geo_key(g["book"]) == gk][0]["sorted"] = True
        for d in geo:
            n += 1 if d["book"].subm else 3
            n += 2 if d["sorted"] else 0
            if self.materialize_pairs is True:
                n += 3 if d["book"].subm else 4
        return n

    def __call__(self, voxel_features, voxel_coords, batch_size):
        while True:
            a = self.launch(voxel_features, voxel_coords, batch_size)
            try:
                return self.collect(a, voxel_coords, batch_size)[0]
            except CapacityOverflow:
                if not self.grow():
                    raise


_TC_FLAG = None


def _tc_available():
    """True once the tcgen05 kernels are linked into the library (pack_weight_bytes answers non-zero)."""
    global _TC_FLAG
    if _TC_FLAG is None:
        _TC_FLAG = _lib.load().fv2p_pack_weight_bytes(27, 64, 64, _lib.MODE_BF16_TC) > 0
    return _TC_FLAG
