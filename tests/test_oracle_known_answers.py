"""Known-answer vectors transcribed from the reference's in-file Rust unit tests and an
independent consensus oracle (tests/test_svar2_reconstruct.py:64-90) -- run against the
CPU oracle so every branch the reference tests is pinned even without cargo."""
import numpy as np

from oracle import oracle as O

N = ord("N")


def _single(v_idxs, v_starts, ilens, shift, alt, alt_off, ref, ref_start, out_len, pad, keep=None):
    """Harness = `run` helper of src/reconstruct/mod.rs:846-900 (one query, one hap)."""
    n = len(v_idxs)
    out = np.full(out_len, pad, np.uint8)
    av, ap = np.zeros(out_len, np.int32), np.zeros(out_len, np.int32)
    O.reconstruct_haplotypes_from_sparse(
        out, np.array([0, out_len]), np.array([[0, ref_start, ref_start + out_len]], np.int32),
        np.array([[shift]], np.int32), np.array([[0]]), np.array([[0], [n]]), np.array(v_idxs, np.int32),
        np.array(v_starts, np.int32), np.array(ilens, np.int32), np.array(alt, np.uint8), np.array(alt_off, np.int64),
        np.array(ref, np.uint8), np.array([0, len(ref)]), pad,
        None if keep is None else np.array(keep, bool), None if keep is None else np.array([0, n]), av, ap)
    return out.tolist(), av.tolist(), ap.tolist()


def test_rust_unit_vectors_reconstruct():
    # src/reconstruct/mod.rs:906-925 no variants
    assert _single([], [], [], 0, [], [0], [10, 20, 30, 40, 50], 1, 3, 0)[0] == [20, 30, 40]
    # :931-954 negative ref_start -> leading pad, annot_ref_pos -1
    out, av, ap = _single([], [], [], 0, [], [0], [1, 2, 3, 4, 5], -2, 5, 9)
    assert out == [9, 9, 1, 2, 3] and av[:2] == [-1, -1] and ap == [-1, -1, 0, 1, 2]
    # :964-988 single SNP
    out, av, _ = _single([0], [2], [0], 0, [84], [0, 1], [65, 67, 71, 84, 65], 0, 5, 0)
    assert out == [65, 67, 84, 84, 65] and av == [-1, -1, 0, -1, -1]
    # :1000ff 2bp insertion truncated at the window end
    assert _single([0], [2], [2], 0, [10, 11, 12], [0, 3], [1, 2, 3, 4, 5], 0, 5, 0)[0] == [1, 2, 10, 11, 12]
    # deletion: pos 1, ilen -2 -> ref[0..1] + anchor + ref[4..]; right pad with INT32_MAX annotation
    out, av, ap = _single([0], [1], [-2], 0, [7], [0, 1], [1, 2, 3, 4, 5], 0, 5, 9)
    assert out == [1, 7, 5, 9, 9]
    assert av == [-1, 0, -1, -1, -1] and ap == [0, 1, 4, 2**31 - 1, 2**31 - 1]
    # overlapping variants at one position: first wins (:108-110)
    out, av, _ = _single([0, 1], [2, 2], [0, 0], 0, [50, 60], [0, 1, 2], [1, 2, 3, 4, 5], 0, 5, 0)
    assert out == [1, 2, 50, 4, 5] and av == [-1, -1, 0, -1, -1]
    # DEL spanning the window start (:99-102, `>=`): pos 0, ilen -2 ends at 3 == ref_start
    assert _single([0], [0], [-2], 0, [1], [0, 1], [1, 2, 3, 4, 5, 6], 3, 3, 0)[0] == [4, 5, 6]
    # keep mask drops the SNP
    assert _single([0], [2], [0], 0, [84], [0, 1], [65, 67, 71, 84, 65], 0, 5, 0, keep=[False])[0] == [65, 67, 71, 84, 65]
    # deletion pushes ref_idx past the contig end -> whole tail is pad (:1088-1105)
    assert _single([0], [3], [-5], 0, [7], [0, 1], [1, 2, 3, 4, 5], 0, 6, 9)[0] == [1, 2, 3, 7, 9, 9]


def test_shift_branches():
    ref = list(range(1, 11))
    # no variants: shift simply advances the reference (:200-205)
    assert _single([], [], [], 3, [], [0], ref, 0, 4, 0)[0] == [4, 5, 6, 7]
    # branch 2 (:123-128): variant at 6, shift 2 -> start at ref[2]
    assert _single([0], [6], [0], 2, [99], [0, 1], ref, 0, 6, 0)[0] == [3, 4, 5, 6, 99, 8]
    # branch 3 (:130-145): insertion at 1 with ALT len 4, shift 3 trims 2 ALT bytes
    assert _single([0], [1], [3], 3, [20, 21, 22, 23], [0, 4], ref, 0, 5, 0)[0] == [22, 23, 3, 4, 5]
    # branch 3 exact consumption (:135-140): ALT fully shifted away
    assert _single([0], [1], [3], 5, [20, 21, 22, 23], [0, 4], ref, 0, 4, 0)[0] == [3, 4, 5, 6]
    # branch 1 (:118-121): variant skipped WITHOUT moving ref_idx
    assert _single([0], [0], [0], 4, [99], [0, 1], ref, 0, 4, 0)[0] == [5, 6, 7, 8]


def test_rc_exhaustive_and_rows():
    # src/reverse.rs:92-104 COMP table incl. pass-through bytes; all 256 byte values
    data = np.arange(256, dtype=np.uint8)
    O.rc_flat_rows_inplace(data, np.array([0, 256]), np.array([True]))
    exp = np.arange(256, dtype=np.uint8)[::-1].copy()
    for a, b in zip(b"ACGT", b"TGCA"):
        exp[255 - a] = b
    assert (data == exp).all()
    d = np.frombuffer(b"ACGTAACG", np.uint8).copy()  # :107-116
    O.rc_flat_rows_inplace(d, np.array([0, 4, 8]), np.array([True, False]))
    assert d.tobytes() == b"ACGTAACG"
    d = np.frombuffer(b"ACN", np.uint8).copy()  # :119-126
    O.rc_flat_rows_inplace(d, np.array([0, 3]), np.array([True]))
    assert d.tobytes() == b"NGT"
    f = np.array([1, 2, 3, 9], np.float32)  # :129-136
    O.reverse_flat_rows_inplace(f, np.array([0, 3, 4]), np.array([True, False]))
    assert f.tolist() == [3, 2, 1, 9]


def test_get_diffs_boundary_rules():
    # src/genotypes/mod.rs:57-84: DEL starting before the window, DEL running past the end
    goi, go = np.array([[0]]), np.array([[0], [2]])
    v_starts, ilens = np.array([8, 18], np.int32), np.array([-5, -6], np.int32)
    d = O.get_diffs_sparse(goi, np.array([0, 1], np.int32), go, ilens, None, None, np.array([10], np.int32),
                           np.array([20], np.int32), v_starts)
    # DEL@8 covers 9..13: clipped to 10..13 -> -4 (+max(10-8-1,0)=+1); DEL@18 covers 19..24 -> -1 (+5)
    assert d.tolist() == [[-5]]


def _svar2_inputs():
    """One query, two haps; vk + dense channels with a tie at pos 20 and a pure DEL."""
    ref = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT", np.uint8)
    # decoded key table: 0 = SNP 'T', 1 = SNP 'G', 2 = pure DEL -2 (empty ALT), 3 = INS "CAA", 4 = SNP 'A'
    key_ilen = np.array([0, 0, -2, 2, 0], np.int32)
    key_alt = np.frombuffer(b"TGCAAA", np.uint8)
    key_alt_off = np.array([0, 1, 2, 2, 5, 6], np.int64)
    vk_pos = np.array([10, 20, 5], np.int32)       # hap0: 10,20 ; hap1: 5
    vk_key = np.array([0, 1, 3], np.int32)
    vk_off = np.array([0, 2, 3], np.int64)
    dense_pos = np.array([15, 20, 30], np.int32)
    dense_key = np.array([2, 4, 0], np.int32)
    dense_range = np.array([[0, 3]], np.int32)
    # hap0 present bits 1,1,1 ; hap1 present bits 0,1,0  -> bitstream LSB-first: 1 1 1 0 1 0
    dense_present = np.array([0b010111], np.uint8)
    dense_present_off = np.array([0, 3, 6], np.int64)
    regions = np.array([[0, 0, 40]], np.int32)
    return dict(ref=ref, key_ilen=key_ilen, key_alt=key_alt, key_alt_off=key_alt_off, vk_pos=vk_pos, vk_key=vk_key,
                vk_off=vk_off, dense_pos=dense_pos, dense_key=dense_key, dense_range=dense_range,
                dense_present=dense_present, dense_present_off=dense_present_off, regions=regions)


def _consensus(ref: bytes, pos, ilen, alleles, q_start, q_end) -> bytes:
    """Independent reconstruction, restated from tests/test_svar2_reconstruct.py:64-90."""
    order = np.argsort(pos, kind="stable")
    out, ref_idx = bytearray(), q_start
    for i in order:
        p, il, al = int(pos[i]), int(ilen[i]), bytes(alleles[i])
        v_end = p - min(0, il) + 1
        if il < 0 and p < q_start and v_end >= q_start:
            ref_idx = v_end
            continue
        if p < ref_idx:
            continue
        if p >= q_end:
            break
        out += ref[ref_idx:p]
        out += al if len(al) > 0 else ref[p:p + 1]
        ref_idx = v_end
    out += ref[ref_idx:q_end]
    return bytes(out)


def test_svar2_merge_order_diffs_and_anchor():
    s = _svar2_inputs()
    diffs = O.hap_diffs_svar2(s["regions"], 2, s["vk_pos"], s["vk_key"], s["vk_off"], s["dense_pos"], s["dense_key"],
                              s["dense_range"], s["dense_present"], s["dense_present_off"], s["key_ilen"])
    # hap0: SNP@10, DEL-2@15, SNP@20 (vk first; dense SNP@20 skipped as overlap), SNP@30 -> -2
    # hap1: INS+2@5, dense SNP@20 -> +2
    assert diffs.tolist() == [[-2, 2]]
    lens = 40 + diffs.ravel().astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    bounds = np.stack([off[:-1], off[1:]], 1)
    out = np.zeros(int(off[-1]), np.uint8)
    O.reconstruct_haplotypes_from_svar2(out, bounds, s["regions"], np.zeros((1, 2), np.int32), s["vk_pos"],
                                        s["vk_key"], s["vk_off"], s["dense_pos"], s["dense_key"], s["dense_range"],
                                        s["dense_present"], s["dense_present_off"], s["key_ilen"], s["key_alt"],
                                        s["key_alt_off"], s["ref"], np.array([0, 40]), N)
    ref = s["ref"].tobytes()
    alle = lambda k: s["key_alt"][s["key_alt_off"][k]:s["key_alt_off"][k + 1]].tobytes()
    # merged order for hap0 (src/svar2/mod.rs:616-649): (10,k0) (15,k2) (20,k1 vk) (20,k4 dense) (30,k0)
    h0 = _consensus(ref, [10, 15, 20, 20, 30], [0, -2, 0, 0, 0], [alle(0), alle(2), alle(1), alle(4), alle(0)], 0, 40)
    h1 = _consensus(ref, [5, 20], [2, 0], [alle(3), alle(4)], 0, 40)
    assert out[off[0]:off[1]].tobytes() == h0
    assert out[off[1]:off[2]].tobytes() == h1
    assert h0[20 - 2] == ord("G")  # the vk key (SNP 'G') won the tie at pos 20, after the 2-bp deletion


def test_fused_ragged_offsets_and_rc():
    # src/ffi/mod.rs:794-811: ragged length = max(ref_len + diff, 0); :842-853 per-row RC
    ref = np.frombuffer(b"ACGTACGTAC", np.uint8)
    v_starts, ilens = np.array([2, 5], np.int32), np.array([2, -1], np.int32)
    alt, alt_off = np.frombuffer(b"GTTA", np.uint8), np.array([0, 3, 4])
    regions = np.array([[0, 0, 8]], np.int32)
    goi, go = np.array([[0, 1]]), np.array([[0, 2], [2, 2]])
    gv = np.array([0, 1], np.int32)
    out, oo = O.reconstruct_haplotypes_fused(regions, np.zeros((1, 2), np.int32), goi, go, gv, v_starts, ilens, alt,
                                             alt_off, ref, np.array([0, 10]), N, -1, None, None,
                                             np.array([False, True]))
    assert oo.tolist() == [0, 9, 17]
    assert out[:9].tobytes() == b"ACGTTTAAT"          # AC + GTT + TA + A (ALT of DEL@5) + T
    assert out[9:].tobytes() == b"ACGTACGT"[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))
    out, av, ap, oo = O.reconstruct_annotated_haplotypes_fused(
        regions, np.zeros((1, 2), np.int32), goi, go, gv, v_starts, ilens, alt, alt_off, ref, np.array([0, 10]), N, 6,
        None, None, np.array([True, False]))
    assert oo.tolist() == [0, 6, 12]
    assert out[:6].tobytes() == b"AAACGT"  # RC of "ACGTTT"
    assert av[:6].tolist() == [-1, -1, 0, 0, 0, -1][::-1]
    assert ap[:6].tolist() == [0, 1, 2, 2, 2, 3][::-1]


def test_onehot_definition():
    h = np.frombuffer(b"ACGTNacgt\x00", np.uint8)
    oh = O.onehot(h)
    assert oh.shape == (10, 4)
    assert oh[:4].tolist() == np.eye(4, dtype=np.uint8).tolist()
    assert not oh[4:].any()


def test_ragged_to_padded():
    data, off = np.arange(7, dtype=np.uint8), np.array([0, 3, 3, 7])
    out = np.full((3, 4), N, np.uint8)
    O.ragged_to_padded(data, off, out, 1, 4)
    assert out.tolist() == [[0, 1, 2, N], [N] * 4, [3, 4, 5, 6]]
