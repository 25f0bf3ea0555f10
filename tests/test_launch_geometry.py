"""Host-side model of how the kernels partition a row among warps, lanes and the edge kernels (cfb_forward.cu lane_setup /
k_fwd_plane_edge, cfb_inverse.cu writer rule / k_inv_plane_edge): for EVERY width the library accepts, each output column is
produced exactly once, border columns get the border filter exactly once, and every halo word a lane reads lies inside the
row.  The GPU parity tests exercise a few dozen widths; this covers all of them without a GPU."""
import pytest

K_STRIP_IN, K_INV_STRIP = 256, 120


def forward_cover(width):
    """-> (columns written by the main kernel, by the edge kernel, lanes flagged right_border, halo reads)"""
    ow = width // 2
    main, halos, right_border = [], [], []
    nstrips = (width + K_STRIP_IN - 1) // K_STRIP_IN
    for strip in range(nstrips):
        for lane in range(32):
            col0 = strip * K_STRIP_IN + lane * 8
            if col0 + 8 > width:
                continue                                    # inactive: not all 8 input columns exist
            main += list(range(col0 // 2, col0 // 2 + 4))
            if col0 + 8 == width:
                right_border.append(col0 // 2 + 3)
            use_lh = lane == 0 and strip > 0
            use_rh = (col0 + 8 < width) and (lane == 31 or col0 + 16 > width)
            if use_lh:
                halos.append((col0 - 2, col0 - 1))
            if use_rh:
                halos.append((col0 + 8, col0 + 9))
            if not (use_rh or col0 + 8 == width):
                assert lane < 31 and col0 + 16 <= width      # right neighbour value comes from an ACTIVE lane
    edge = list(range((width // 8) * 4, ow))
    return main, edge, right_border, halos


@pytest.mark.parametrize("width", list(range(16, 4200, 2)))
def test_forward_row_partition(width):
    ow = width // 2
    main, edge, right_border, halos = forward_cover(width)
    assert sorted(main + edge) == list(range(ow))            # every output column exactly once
    assert len(edge) <= 3
    if width % 8 == 0:
        assert right_border == [ow - 1] and not edge          # the last full lane applies the border filter
    else:
        assert not right_border and edge[-1] == ow - 1        # the edge kernel owns the right border column
    for a, b in halos:
        assert 0 <= a and b < width                           # halo samples exist


def inverse_cover(bw):
    main = []
    nstrips = (bw + K_INV_STRIP - 1) // K_INV_STRIP
    right_border = []
    for strip in range(nstrips):
        for lane in range(32):
            col0 = strip * K_INV_STRIP - 4 + lane * 4
            active = 0 <= col0 < bw
            writer = active and 1 <= lane <= 30 and col0 + 4 <= bw
            if writer:
                main += list(range(col0, col0 + 4))
                if col0 + 4 == bw:
                    right_border.append(col0 + 3)
                else:
                    # its right tap is band column col0 + 4, loaded by the next lane (halo lane 31 included)
                    assert col0 + 4 < bw
            if active:
                # a partial lane loads 8 bytes from col0: must stay inside the band pitch ALIGN16(2 * bw)
                assert 2 * col0 + 8 <= (2 * bw + 15) // 16 * 16
    edge = list(range((bw // 4) * 4, bw))
    return main, edge, right_border


@pytest.mark.parametrize("bw", list(range(6, 2100)))
def test_inverse_row_partition(bw):
    main, edge, right_border = inverse_cover(bw)
    assert sorted(main + edge) == list(range(bw))
    assert len(edge) <= 3
    if bw % 4 == 0:
        assert right_border == [bw - 1] and not edge
    else:
        assert not right_border and edge[-1] == bw - 1
