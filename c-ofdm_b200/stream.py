"""Sharding the streaming receiver (rx.cpp's acquisition loop, cofdm_rx_stream) over ranks / GPUs.

The loop is a sequential state machine: where it looks next depends on the frame it found last
(rx.cpp:126-198).  Two chains that start from different states nevertheless coincide from the first
frame both of them detect (after a detection the state is `preamble position + message.size`, whatever
came before).  So a long capture is cut into contiguous ranges of whole SDR blocks; every rank runs the
unmodified loop on its range PLUS one extra block, starting cold; where two neighbouring ranks overlap,
the earlier rank's chain (the true one) is followed until it reaches a preamble position the later rank
also found, and the later rank's list is used from there on.  No data-path communication: each rank
reads its own slice; the lists (a few bytes per frame) are gathered at the end.
"""
import numpy as np

from .dist import shard_range


def block_samples(sizes):
    """samples per SDR block: output_size * rx_buf_size (sdr.hpp:141)"""
    return sizes.output_size * sizes.rx_buf_size


def shard_slice(n_samples, sizes, rank, world, overlap_blocks=1):
    """-> (first_sample, last_sample_exclusive, first_block, last_block_exclusive) of `rank`'s slice of a capture,
    including `overlap_blocks` extra blocks at the end (except for the last rank)."""
    blk = block_samples(sizes)
    n_blocks = n_samples // blk
    b0, b1 = shard_range(n_blocks, rank, world)
    b1x = min(n_blocks, b1 + (overlap_blocks if b1 < n_blocks else 0))
    return b0 * blk, b1x * blk, b0, b1


def merge_shards(shards, sizes):
    """shards: list over ranks (in order) of (positions [absolute sample indices], bytes [n, usefull_size],
    first_block, last_block_exclusive).  Returns (positions, bytes) of the whole capture, identical to one
    sequential pass, plus the number of boundaries that did not re-synchronise inside the overlap (should be 0).
    `bytes` may be any per-frame rows (payloads, or tags telling which rank holds the payload)."""
    blk = block_samples(sizes)
    out_pos, out_rows, unmerged = [], [], 0
    carry_pos, carry_rows = np.zeros(0, np.int64), None          # the previous rank's frames beyond its own range
    width, dtype = sizes.usefull_size, np.uint8
    for pos, by, b0, b1 in shards:
        pos = np.asarray(pos, dtype=np.int64)
        by = np.asarray(by)
        if by.ndim == 2:
            width, dtype = by.shape[1], by.dtype
        start = 0
        if len(carry_pos):
            # follow the previous (true) chain until it meets this rank's chain
            # (positions are increasing: a binary search of the short carry list in this rank's list, no sort)
            at = np.searchsorted(pos, carry_pos)
            hit = np.nonzero((at < len(pos)) & (pos[np.minimum(at, max(len(pos) - 1, 0))] == carry_pos))[0] if len(pos) else np.zeros(0, np.int64)
            if len(hit):
                k = int(hit[0])
                out_pos.append(carry_pos[:k])
                out_rows.append(carry_rows[:k])
                start = int(np.searchsorted(pos, carry_pos[k]))
            else:
                unmerged += 1
                out_pos.append(carry_pos)
                out_rows.append(carry_rows)
                start = int(np.searchsorted(pos, carry_pos[-1] + 1))
        own_end = b1 * blk
        # a frame is owned by the range that contains its preamble; the rest of the list is carried over
        n_own = max(start, int(np.searchsorted(pos, own_end)))    # positions are increasing
        out_pos.append(pos[start:n_own])
        out_rows.append(by[start:n_own])
        carry_pos, carry_rows = pos[n_own:], by[n_own:]
    out_pos.append(carry_pos)
    if carry_rows is not None:
        out_rows.append(carry_rows)
    rows = [r for r in out_rows if len(r)]
    b = np.concatenate(rows) if rows else np.zeros((0, width), dtype)
    return np.concatenate(out_pos) if out_pos else np.zeros(0, np.int64), b, unmerged


def merge_ranges(lists, sizes):
    """The merge of merge_shards without per-frame rows: lists = [(positions, b0, b1)] over ranks in order.  The frames of
    rank r that belong to the merged list are ONE contiguous slice [lo_r, hi_r) of its own list (its own range from where the
    previous rank's chain met it, plus the first frames of its overlap run up to where the next rank's chain meets it).
    Returns (merged positions, [(lo_r, hi_r)], unmerged boundaries)."""
    blk = block_samples(sizes)
    ranges, unmerged = [], 0
    prev = None                                   # (pos, n_own) of the previous rank
    starts = []
    for pos, b0, b1 in lists:
        pos = np.asarray(pos, dtype=np.int64)
        start = 0
        if prev is not None:
            ppos, pn_own = prev
            carry = ppos[pn_own:]
            k = len(carry)                        # frames of the previous rank's overlap run that stay
            if len(carry):
                at = np.searchsorted(pos, carry)
                hit = np.nonzero((at < len(pos)) & (pos[np.minimum(at, max(len(pos) - 1, 0))] == carry))[0] if len(pos) else np.zeros(0, np.int64)
                if len(hit):
                    k = int(hit[0])
                    start = int(at[k])
                else:
                    unmerged += 1
                    start = int(np.searchsorted(pos, carry[-1] + 1))
            ranges[-1] = (ranges[-1][0], pn_own + k)
        n_own = max(start, int(np.searchsorted(pos, b1 * blk)))
        ranges.append((start, n_own))
        prev = (pos, n_own)
    if prev is not None:
        ranges[-1] = (ranges[-1][0], len(prev[0]))     # the last rank keeps everything it found
    mpos = np.concatenate([np.asarray(p, dtype=np.int64)[lo:hi] for (p, _, _), (lo, hi) in zip(lists, ranges)]) if lists else np.zeros(0, np.int64)
    return mpos, ranges, unmerged


def gather_positions(pos_abs, b0, b1, max_frames, device=None):
    """ONE fixed-size collective: every rank contributes [count, b0, b1, positions padded to max_frames] (int64) and gets
    everybody's back.  max_frames = an upper bound known to all ranks (samples per slice / message.size + 2)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    pos_abs = np.asarray(pos_abs, dtype=np.int64)
    rec = np.full(max_frames + 3, -1, dtype=np.int64)
    rec[0], rec[1], rec[2] = len(pos_abs), b0, b1
    rec[3:3 + len(pos_abs)] = pos_abs
    mine = torch.from_numpy(rec).to(device)
    allr = torch.empty(world * (max_frames + 3), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allr, mine)
    allr = allr.cpu().numpy().reshape(world, max_frames + 3)
    return [(allr[r, 3:3 + int(allr[r, 0])], int(allr[r, 1]), int(allr[r, 2])) for r in range(world)]


def rx_stream_sharded(run, capture_i16, sizes, world):
    """run(capture_slice) -> (positions, bytes).  Emulates `world` ranks in one process (tests, single GPU)."""
    n = capture_i16.shape[0]
    shards = []
    for r in range(world):
        s0, s1, b0, b1 = shard_slice(n, sizes, r, world)
        pos, by = run(capture_i16[s0:s1])
        shards.append((np.asarray(pos) + s0, by, b0, b1))
    return merge_shards(shards, sizes)


def gather_frame_lists(pos_abs, b0, b1, device=None):
    """The one exchange of the sharded stream receiver: every rank contributes its list of absolute preamble positions
    (int64) and its block range; every rank gets all of them back.  Two fixed-size collectives (all_gather of the counts,
    all_gather of the lists padded to the longest) on the process group's own backend -- NCCL over NVLink on GPUs, gloo on
    CPU -- no pickling, no payload bytes.  Returns a list over ranks of (positions, tags, b0, b1), tags[i] = (rank, i)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    pos_abs = np.asarray(pos_abs, dtype=np.int64)
    head = torch.tensor([len(pos_abs), b0, b1], dtype=torch.int64, device=device)
    heads = torch.empty(world * 3, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(heads, head)
    heads = heads.cpu().numpy().reshape(world, 3)
    width = max(1, int(heads[:, 0].max()))
    mine = torch.full((width,), -1, dtype=torch.int64, device=device)
    if len(pos_abs):
        mine[: len(pos_abs)] = torch.from_numpy(pos_abs).to(device)
    allp = torch.empty(world * width, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allp, mine)
    allp = allp.cpu().numpy().reshape(world, width)
    out = []
    for r in range(world):
        n = int(heads[r, 0])
        tag = np.stack([np.full(n, r, np.int64), np.arange(n, dtype=np.int64)], axis=1)
        out.append((allp[r, :n].copy(), tag, int(heads[r, 1]), int(heads[r, 2])))
    return out


def rx_stream_distributed(modem, capture_i16, shards=1):
    """torchrun entry: every rank receives its slice of the capture (a numpy array, or a torch tensor already on
    its GPU) on its own GPU -- itself cut into `shards` ranges scanned concurrently.  The only communication is
    gather_frame_lists (positions, a few bytes per frame); every rank then merges the chains itself and learns which
    of ITS frames belong to the merged list.  Payloads never leave the rank that decoded them.
    Returns (positions_of_the_whole_capture, tags [n, 2] = (rank, index in that rank's list), my_bytes, unmerged)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    s = modem.sizes
    s0, s1, b0, b1 = shard_slice(capture_i16.shape[0], s, rank, world)
    pos, by = modem.rx_stream(capture_i16[s0:s1], shards=shards)
    pos_abs = np.asarray(pos, dtype=np.int64) + s0
    if world == 1:
        tag = np.stack([np.zeros(len(pos_abs), np.int64), np.arange(len(pos_abs), dtype=np.int64)], axis=1)
        lists = [(pos_abs, tag, b0, b1)]
    else:
        lists = gather_frame_lists(pos_abs, b0, b1)
    mpos, mtag, unmerged = merge_shards(lists, s)
    return mpos, mtag, by, unmerged
