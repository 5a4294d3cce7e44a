"""Frame-major row order of the temporal stage (csrc/traj_fused.cuh `traj_row_canonical`, csrc/attn_tc.cuh epilogue; pair-mode bit 16).

CPU restatement of the two index maps the kernels use, checked for the properties the design relies on: every 128-row tile holds tokens
of ONE frame (so the tile's x_t image is its x_diag, WC/temporal_attention.py:61-63), the map from valid tile-order rows to canonical
tokens is a bijection, and the attention kernel's output row of (sequence, query) is the row the temporal kernel reads it from."""
import numpy as np
import pytest


def pass_to_canonical(p, mode, B, T, H, W):
    if mode == 1:                      # H pass: p = ((b*W + w)*T + t)*H + h
        h = p % H; r = p // H; t = r % T; r //= T; w = r % W; b = r // W
    elif mode == 2:                    # W pass: p = ((b*H + h)*T + t)*W + w
        w = p % W; r = p // W; t = r % T; r //= T; h = r % H; b = r // H
    else:
        return p
    return ((b * T + t) * H + h) * W + w


def traj_row_canonical(r, mode, B, T, H, W, rpad, rt, n):
    t, rem = divmod(r, rpad)
    if t >= T or rem >= rt:
        return -1
    seq, j = divmod(rem, n)
    return pass_to_canonical((seq * T + t) * n + j, mode, B, T, H, W)


def attn_out_row(seq, qi, n, rpad, magic):
    t = (qi * magic) >> 32             # the kernel's reciprocal multiply
    assert t == qi // n
    return t * rpad + seq * n + (qi - t * n)


@pytest.mark.parametrize("B,T,H,W,mode", [(3, 2, 21, 21, 1), (2, 2, 41, 41, 2), (2, 5, 15, 20, 1), (2, 5, 15, 20, 2), (1, 2, 5, 5, 0), (9, 1, 7, 3, 2)])
def test_frame_major_rows(B, T, H, W, mode):
    if mode == 1:
        num_seq, n = B * W, H
    elif mode == 2:
        num_seq, n = B * H, W
    else:
        num_seq, n = B, H * W
    rt = num_seq * n
    rpad = (rt + 127) // 128 * 128
    rows = T * rpad
    canon = np.array([traj_row_canonical(r, mode, B, T, H, W, rpad, rt, n) for r in range(rows)])
    valid = canon >= 0
    assert valid.sum() == B * T * H * W
    assert sorted(canon[valid].tolist()) == list(range(B * T * H * W))            # bijection onto the tokens
    frame_of = (canon // (H * W)) % T
    for tile in range(rows // 128):                                               # one frame per tile: x_diag(tile) = x_t(tile)
        sl = slice(tile * 128, tile * 128 + 128)
        ts = set(frame_of[sl][valid[sl]].tolist())
        assert ts <= {tile // (rpad // 128)}
    magic = (1 << 32) // n + 1
    N = T * n
    for seq in range(num_seq):                                                    # the attention kernel writes where the temporal kernel reads
        for qi in range(N):
            r = attn_out_row(seq, qi, n, rpad, magic)
            assert canon[r] == pass_to_canonical(seq * N + qi, mode, B, T, H, W)
