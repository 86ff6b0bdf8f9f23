"""Index arithmetic of the (experimental, SERL_EPI_COAL=1) coalesced conv epilogues, replayed on the CPU: every
(row, channel-chunk) written to the per-warp shared transpose tile must come back at the lane that stores it to the right
global address, and neither the writes nor the reads may bank-conflict (16-byte slots of a 128-byte line, 8 lanes per wavefront).
Mirrors serl_b200/csrc/conv3x3_tcgen05.cu (4 KB tiles, 64-channel groups) and conv_tcgen05.cu (2 KB tiles, 32-channel groups)."""
import pytest


def _c3_write(lane, ch):
    return lane * 128 + ((ch ^ (lane & 7)) << 4)


def _c3_read(lane, i):
    r, cc = (lane >> 3) + 4 * i, lane & 7
    return r, cc, r * 128 + ((cc ^ (r & 7)) << 4)


def _ctc_write(lane, ch):
    return lane * 64 + ((ch ^ ((lane >> 1) & 3)) << 4)


def _ctc_read(lane, i):
    r, cc = (lane >> 2) + 8 * i, lane & 3
    return r, cc, r * 64 + ((cc ^ ((r >> 1) & 3)) << 4)


@pytest.mark.parametrize("write,read,chunks,iters", [(_c3_write, _c3_read, 8, 8), (_ctc_write, _ctc_read, 4, 4)])
def test_transpose_tile_round_trip_and_conflicts(write, read, chunks, iters):
    tile = {}
    for lane in range(32):
        for ch in range(chunks):
            off = write(lane, ch)
            assert off not in tile
            tile[off] = (lane, ch)
    seen = set()
    for i in range(iters):
        for lane in range(32):
            r, cc, off = read(lane, i)
            assert tile[off] == (r, cc)                      # lane stores chunk cc of row r: 8 (4) lanes cover a row's 128 (64) bytes
            seen.add((r, cc))
    assert len(seen) == 32 * chunks
    slot = lambda off: (off // 16) % 8
    for ch in range(chunks):                                 # writes: all lanes store the same chunk index of their own row
        for q in range(4):
            s = [slot(write(l, ch)) for l in range(q * 8, q * 8 + 8)]
            assert len(set(s)) == 8
    for i in range(iters):
        for q in range(4):
            s = [slot(read(l, i)[2]) for l in range(q * 8, q * 8 + 8)]
            assert len(set(s)) == 8
