"""Host logic of the work list that hydro_stage_vec runs (amx_vec_common.cuh pick_shares + the segment loop at the top
of the kernel), transcribed to Python: every plane of every tile is covered exactly once, shares are balanced, and
the launch never exceeds the number of resident blocks.  (The kernels themselves are held to the generic family bit
for bit on the GPU, with AMX_SHARE_SKEW forcing ragged shares.)"""
import pytest


def pick_shares(nx, tiles, blocks_per_sm, skew=0):
    slots = 148 * blocks_per_sm
    total = tiles * nx
    nb = (total // 16 if total // 16 > 0 else 1) if total // 16 < slots else slots
    w = (total + nb - 1) // nb
    nch = max(1, (nx + w // 2) // w)
    xchunk = (nx + nch - 1) // nch
    nch = (nx + xchunk - 1) // xchunk
    padded = nch * tiles * xchunk
    w = (padded + nb - 1) // nb + skew
    return (padded + w - 1) // w, xchunk, w


def segments(nx, tiles, xchunk, work, block):
    """the kernel's loop: [(tile, chunk, xs, xe)] of one block"""
    nchunks = (nx + xchunk - 1) // xchunk
    u, u_end = block * work, min((block + 1) * work, nchunks * tiles * xchunk)
    out = []
    while u < u_end:
        tl = u // xchunk
        xoff = u - tl * xchunk
        chunk, tile = tl // tiles, tl % tiles
        xs = chunk * xchunk + xoff
        ce = min(chunk * xchunk + xchunk, nx)
        xe = min(ce, xs + (u_end - u))
        u = (tl + 1) * xchunk if xe >= ce else u_end
        if xs < xe:
            out.append((tile, chunk, xs, xe))
    return out


@pytest.mark.parametrize("skew", [0, 5, 7])
@pytest.mark.parametrize("nx,tiles", [(256, 152), (512, 592), (128, 2368), (64, 10), (33, 3), (8, 1), (100, 7),
                                      (48, 8), (32, 6), (40, 6), (272, 1), (1024, 3), (250, 19)])
def test_every_plane_tile_once_and_balanced(nx, tiles, skew):
    nb, xchunk, work = pick_shares(nx, tiles, 2, skew)
    assert 1 <= nb <= 296
    seen, cost = set(), []
    for b in range(nb):
        segs = segments(nx, tiles, xchunk, work, b)
        cost.append(sum(xe - xs for _, _, xs, xe in segs))
        for tile, chunk, xs, xe in segs:
            assert chunk * xchunk <= xs < xe <= min((chunk + 1) * xchunk, nx)
            for x in range(xs, xe):
                assert (tile, x) not in seen
                seen.add((tile, x))
    assert len(seen) == tiles * nx
    assert max(cost) <= work
    if skew == 0 and tiles * nx >= 296 * 16:
        assert max(cost) - min(c for c in cost if c) <= max(xchunk, work) and max(cost) * nb < 1.15 * tiles * nx + nb * 16


def test_benchmark_geometry():
    """256^3: 8 x 19 tiles of 14 x 32 columns -> 295 blocks of 132 plane-tiles in two x chunks of 128 planes"""
    assert pick_shares(256, 152, 2) == (295, 128, 132)
    segs = [len(segments(256, 152, 128, 132, b)) for b in range(295)]
    assert max(segs) <= 3 and sum(segs) / len(segs) < 2.2
