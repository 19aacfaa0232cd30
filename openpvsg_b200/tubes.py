"""Tube linking of per-frame VPS results (reference models/mask2former_vps/utils.py:20-89
``concat_seq``) and its multi-GPU form.

Frames of a clip are independent work items, so a clip is sharded over ranks in contiguous
blocks; the only exchange of the whole path is here: an all-gather of the kept
(segment id, query feature) entries of every frame.  Masks are never gathered -- each rank
RLE-encodes the frames it owns.  Wire formats follow the reference:
``masks.txt`` lines ``frame id cid h w rle`` (models/unitrack/utils/io.py:14-37) and the
per-tube feature lists of ``query_feats.pickle`` (utils.py:75-89).
"""
import numpy as np
import torch


# -------------------------------------------------------------------------- RLE -------
def rle_counts(mask):
    """COCO RLE run lengths of a [H,W] binary mask (column-major, first run counts zeros)."""
    flat = np.asarray(mask, dtype=np.uint8).reshape(-1, order='F')
    if flat.size == 0:
        return [0]
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    bounds = np.concatenate(([0], change, [flat.size]))
    runs = np.diff(bounds).tolist()
    return runs if flat[0] == 0 else [0] + runs


def rle_string(counts):
    """pycocotools ``rleToString``: 5 bits per char + continuation bit, deltas from the 3rd run."""
    out = []
    for i, x in enumerate(counts):
        x = int(x)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            c = x & 0x1f
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(chr(c + 48))
    return ''.join(out)


def rle_decode(string, h, w):
    """Inverse of rle_string + rle_counts (pycocotools ``rleFrString`` + decode)."""
    counts, p, m = [], 0, 0
    while p < len(string):
        x, k, more = 0, 0, True
        while more:
            c = ord(string[p]) - 48
            x |= (c & 0x1f) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if m > 2:
            x += counts[m - 2]
        counts.append(x)
        m += 1
    flat = np.zeros(h * w, np.uint8)
    pos, val = 0, 0
    for c in counts:
        flat[pos:pos + c] = val
        pos += c
        val ^= 1
    return flat.reshape((h, w), order='F')


def rle_string_np(counts):
    """Vectorised ``rle_string`` (same bytes): counts int array -> str."""
    c = np.asarray(counts, dtype=np.int64)
    x = c.copy()
    if c.size > 3:
        x[3:] -= c[1:-2]
    n = x.size
    chars = np.zeros((n, 8), np.uint8)
    nchar = np.zeros(n, np.int64)
    active = np.ones(n, bool)
    for r in range(8):
        low = x & 0x1f
        x = x >> 5
        more = np.where((low & 0x10) != 0, x != -1, x != 0)
        ch = (low | (more.astype(np.int64) << 5)) + 48
        chars[active, r] = ch[active]
        nchar[active] += 1
        active &= more
        if not active.any():
            break
    keep = np.arange(8)[None, :] < nchar[:, None]
    return chars[keep].tobytes().decode('ascii')


def rle_from_events(ev_pos, ev_slot, n_events, seg_ids, h, w, native=True):
    """Host side of ``ops.rle_events`` for one frame: the run-boundary events (column-major positions,
    segment slots) -> {segment id: COCO RLE string}, identical to ``rle_string(rle_counts(pan == id))``.
    seg_ids: kept segment ids in slot order (first appearance among the kept rows of seg_info).
    native: use the library's C++ host routine (pvsg_rle_strings_host); the numpy path is kept as its
    cross-check."""
    n = int(n_events)
    if not seg_ids:
        return {}
    if native:
        from . import lib as _l
        lib = _l.load()
        pos = np.ascontiguousarray(ev_pos[:n]).view(np.uint32) if n else np.zeros(1, np.uint32)
        slot = np.ascontiguousarray(ev_slot[:n], dtype=np.int16) if n else np.zeros(1, np.int16)
        cap = 6 * (n + 2 * len(seg_ids)) + 16
        buf = np.empty(cap, np.uint8)
        off = np.empty(len(seg_ids) + 1, np.int64)
        rc = lib.pvsg_rle_strings_host(pos.ctypes.data, slot.ctypes.data, n, len(seg_ids), h * w, buf.ctypes.data, cap,
                                       off.ctypes.data)
        if rc < 0:
            raise _l.PvsgError(f'pvsg_rle_strings_host failed ({rc})')
        raw = buf[:rc].tobytes()
        return {int(sid): raw[off[k]:off[k + 1]].decode('ascii') for k, sid in enumerate(seg_ids)}
    pos = np.asarray(ev_pos[:n]).astype(np.uint32).astype(np.int64)     # stored as the bits of a uint32
    slot = np.asarray(ev_slot[:n]).astype(np.int16)
    order = np.argsort(slot, kind='stable')                              # events of a segment stay in walk order
    pos, slot = pos[order], slot[order]
    bounds = np.searchsorted(slot, np.arange(len(seg_ids) + 1))
    out = {}
    for k, sid in enumerate(seg_ids):
        p = pos[bounds[k]:bounds[k + 1]]
        counts = np.diff(np.concatenate(([0], p, [h * w])))
        out[int(sid)] = rle_string_np(counts)
    return out


def slot_ids(seg_info):
    """Kept segment ids of a seg_info row in slot order (as pvsg_rle_events numbers them)."""
    n = int(seg_info[0])
    ids = []
    for seg in np.asarray(seg_info[1:1 + 4 * n]).reshape(n, 4)[:, 2].tolist():
        if seg >= 0 and seg not in ids:
            ids.append(int(seg))
    return ids


# ------------------------------------------------------------------- tube linking -----
class TubeLinker:
    """Incremental ``concat_seq``: tube id = 1 + order of first appearance of a panoptic id."""

    def __init__(self):
        self.object_list = []
        self._tube_of = {}      # panoptic id -> tube id (the reference searches object_list linearly, utils.py:38-42)
        self.feat_tubes = {}
        self.rows = []          # (frame (1-based), tube id, class id, h, w, rle string)
        self.frame_seg_ids = []  # per frame: kept panoptic ids in slot order (what pvsg_rle_events / pvsg_tube_overlap index)
        self.num_frames = 0

    def _tube(self, ins_id):
        tid = self._tube_of.get(ins_id)
        if tid is None:
            self.object_list.append(ins_id)
            tid = self._tube_of[ins_id] = len(self.object_list)
            self.feat_tubes[tid] = {}
        return tid

    def add_frame(self, seg_ids, feats, pan=None, rle=None, hw=None):
        """seg_ids: iterable of panoptic ids kept in this frame (reference dict order);
        feats: matching [n,256] array; pan: optional int32 [H,W] map (-> masks.txt rows), or
        rle: {segment id: RLE string} from the device encoder together with hw = (H, W)."""
        frame_id = self.num_frames
        seg_ids = [int(i) for i in seg_ids]
        self.frame_seg_ids.append(seg_ids)
        for ins_id, feat in zip(seg_ids, feats):
            tid = self._tube(ins_id)
            self.feat_tubes[tid][frame_id] = dict(query_feat=np.array(feat, dtype=np.float32, copy=True).reshape(-1),
                                                  cls_id=int(ins_id % 1000))
            if rle is not None:
                self.rows.append((frame_id + 1, tid, int(ins_id % 1000), hw[0], hw[1], rle[ins_id]))
            elif pan is not None:
                mask = (pan == ins_id)
                self.rows.append((frame_id + 1, tid, int(ins_id % 1000), mask.shape[0], mask.shape[1],
                                  rle_string(rle_counts(mask))))
        self.num_frames += 1

    def add_frames_bulk(self, counts, ids, feats):
        """Link a whole block of frames from compact arrays (the all-gathered form): counts int [F] kept segments
        per frame, ids int [n] their panoptic ids in frame order, feats fp32 [n,256].  Same result as F calls of
        ``add_frame``; the per-entry Python work is two dict operations (no array copies, no list searches)."""
        counts = np.asarray(counts, np.int64)
        ids = np.asarray(ids, np.int64)
        feats = np.ascontiguousarray(feats, np.float32).reshape(len(ids), -1)
        if int(counts.sum()) != len(ids):
            raise ValueError('add_frames_bulk: counts do not add up to the number of entries')
        frame_of = np.repeat(np.arange(len(counts)), counts) + self.num_frames
        id_list = ids.tolist()
        bounds = np.concatenate(([0], np.cumsum(counts))).tolist()
        for f in range(len(counts)):
            self.frame_seg_ids.append(id_list[bounds[f]:bounds[f + 1]])
        for k, (ins_id, fr) in enumerate(zip(id_list, frame_of.tolist())):
            self.feat_tubes[self._tube(ins_id)][fr] = dict(query_feat=feats[k], cls_id=ins_id % 1000)
        self.num_frames += len(counts)

    def tube_features(self, feature_dim=256):
        """[N_tubes, T, 256] with zero rows for absent frames -- the relation head's input
        (utils/relation_matching.py:431-442, datasets/datasets/pvsg_relation.py:47-53)."""
        out = np.zeros((len(self.object_list), self.num_frames, feature_dim), np.float32)
        for tid, frames in self.feat_tubes.items():
            for f, d in frames.items():
                out[tid - 1, f] = d['query_feat']
        return out

    def frame_tube_ids(self):
        """Per frame: tube id of every slot (slot = position of a kept panoptic id in the frame's result)."""
        return [[self._tube_of[i] for i in ids] for ids in self.frame_seg_ids]

    def masks_txt(self):
        return ''.join(f'{fr} {tid} {cid} {h} {w} {rle}\n' for fr, tid, cid, h, w, rle in self.rows)


def concat_seq(outputs):
    """outputs: list over frames of [result dict] (what single_gpu_test collects).  Returns the
    TubeLinker holding masks.txt rows and the per-tube query features."""
    linker = TubeLinker()
    for output in outputs:
        output = output[0]
        ids = list(output['query_feats'].keys())
        feats = [np.asarray(torch.as_tensor(output['query_feats'][k][0]).cpu()) for k in ids]
        if 'rle' in output:     # FrameRunner(rle=True): strings from the device encoder
            linker.add_frame(ids, feats, rle=output['rle'], hw=output['pan_results'].shape)
        else:
            linker.add_frame(ids, feats, output.get('pan_results'))
    return linker


# ------------------------------------------------------------------ multi-GPU ---------
def shard_frames(num_frames, world_size, rank):
    """Contiguous block of ceil(T / G) frames per rank."""
    per = (num_frames + world_size - 1) // world_size
    lo = min(rank * per, num_frames)
    return lo, min(lo + per, num_frames)


def pack_frames(frame_entries, feature_dim=256):
    """frame_entries: list over local frames of (seg_ids list, feats [n,256]).  Compact form of a frame block:
    (counts int32 [F], ids int32 [n], feats fp32 [n,feature_dim]) -- only kept entries, no padding."""
    counts = np.fromiter((len(sid) for sid, _ in frame_entries), np.int32, len(frame_entries))
    n = int(counts.sum())
    ids = np.fromiter((int(i) for sid, _ in frame_entries for i in sid), np.int32, n)
    feats = np.empty((n, feature_dim), np.float32)
    k = 0
    for sid, ft in frame_entries:
        if len(sid):
            feats[k:k + len(sid)] = np.asarray(ft, np.float32).reshape(len(sid), feature_dim)
            k += len(sid)
    return counts, ids, feats


def gather_and_link(frame_entries, num_frames, max_segments=100, device='cpu', feature_dim=256):
    """The one exchange of the path: all-gather the kept (segment id, query feature) entries of every rank's
    frame block (NCCL over NVLink when ``device`` is a GPU, gloo on CPU) and link tubes over the whole clip.
    Every rank returns the same TubeLinker (without masks.txt rows: masks stay on the rank that owns the frame).

    Wire format per rank: ONE fp32 buffer [per + n_max * (1 + feature_dim)] = frame counts | ids | features
    (ids / counts < 2^24 are exact in fp32), n_max = the largest entry count of any rank (a 4-byte all-gather
    precedes the payload).  Only kept entries travel: ~4 KB per frame instead of the 100-slot padded 103 KB."""
    import torch.distributed as dist
    ws = dist.get_world_size() if dist.is_initialized() else 1
    per = (num_frames + ws - 1) // ws
    if len(frame_entries) > per:
        raise ValueError(f'gather_and_link: {len(frame_entries)} local frames exceed the block size {per}')
    counts, ids, feats = pack_frames(frame_entries, feature_dim)
    if ids.size and (int(ids.max()) >= 1 << 24 or int(ids.min()) < 0 or int(counts.max()) > max_segments):
        raise ValueError('gather_and_link: segment ids must be in [0, 2^24) and at most max_segments per frame')
    if ws == 1:
        linker = TubeLinker()
        linker.add_frames_bulk(counts, ids, feats)
        return linker
    n = torch.tensor([len(ids)], dtype=torch.int64, device=device)
    all_n = [torch.empty_like(n) for _ in range(ws)]
    dist.all_gather(all_n, n)
    all_n = [int(t.item()) for t in all_n]
    n_max = max(all_n)
    buf = np.zeros(per + n_max * (1 + feature_dim), np.float32)
    buf[:len(counts)] = counts
    buf[per:per + len(ids)] = ids
    buf[per + n_max:per + n_max + feats.size] = feats.reshape(-1)
    local = torch.from_numpy(buf).to(device, non_blocking=True)
    gathered = torch.empty(ws * buf.size, dtype=torch.float32, device=device)
    dist.all_gather(list(gathered.view(ws, -1).unbind(0)), local)
    g = gathered.view(ws, -1).cpu().numpy()
    linker = TubeLinker()
    left = num_frames
    for r in range(ws):
        f = min(per, left)
        left -= f
        c = g[r, :f].astype(np.int64)
        m = int(c.sum())
        if m > all_n[r]:
            raise ValueError('gather_and_link: inconsistent frame counts')
        linker.add_frames_bulk(c, g[r, per:per + m].astype(np.int64),
                               g[r, per + n_max:per + n_max + m * feature_dim].reshape(m, feature_dim))
    return linker


def minvis_link_sharded(local_embeds, num_frames, solve_pairs=None, compose=None):
    """MinVIS query matching of a clip whose frames are sharded across ranks (reference
    mask2former_min_vis.py:176-181 + 244-258 run sequentially on one GPU).

    ``local_embeds`` [F_local, Q, C]: the raw query embeddings of this rank's contiguous frame block
    (``shard_frames``).  Exchange 1: all-gather of the blocks (Q*C*4 = 100 KB per frame).  Each rank then solves the
    assignment problems of the frame pairs (t-1, t) with t in ITS block -- the pair that crosses into the previous block
    uses the gathered halo frame -- and exchange 2 all-gathers the matchings (Q int32 per pair).  Every rank returns the
    same perms int64 [T, Q]: position i of the clip-long ordering holds query perms[t][i] of frame t.

    ``solve_pairs(embeds [n+1, Q, C]) -> int32 [n, Q]`` defaults to the device kernels (ops.minvis_chain);
    ``compose(sigma [T-1, Q], Q) -> [T, Q]`` to ops.perm_chain.  The CPU tests inject host versions of both."""
    import torch.distributed as dist
    if solve_pairs is None or compose is None:
        from . import ops
        solve_pairs = solve_pairs or ops.minvis_chain
        compose = compose or ops.perm_chain
    ws = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if ws > 1 else 0
    per = (num_frames + ws - 1) // ws
    lo, hi = shard_frames(num_frames, ws, rank)
    if local_embeds.shape[0] != hi - lo:
        raise ValueError(f'minvis_link_sharded: rank {rank} holds {local_embeds.shape[0]} frames, its block has {hi - lo}')
    Q, C = local_embeds.shape[1:]
    dev = local_embeds.device
    if ws == 1:
        everything = local_embeds
    else:
        pad = torch.zeros(per, Q, C, device=dev, dtype=torch.float32)
        pad[:hi - lo] = local_embeds
        gathered = torch.empty(ws * per, Q, C, device=dev, dtype=torch.float32)
        dist.all_gather(list(gathered.view(ws, per, Q, C).unbind(0)), pad)
        everything = gathered[:num_frames]          # blocks are contiguous and only the last may be short
    first = max(lo, 1)                              # pairs (t-1, t), t in [first, hi)
    sigma_local = torch.zeros(per, Q, device=dev, dtype=torch.int32)
    if hi > first:
        sigma_local[first - lo:hi - lo] = solve_pairs(everything[first - 1:hi].contiguous()).to(torch.int32)
    if ws == 1:
        sigma = sigma_local[1:num_frames]
    else:
        allsig = torch.empty(ws * per, Q, device=dev, dtype=torch.int32)
        dist.all_gather(list(allsig.view(ws, per, Q).unbind(0)), sigma_local)
        sigma = allsig[1:num_frames]                # row t-1 = matching of the pair (t-1, t)
    return compose(sigma.contiguous(), Q).long()
