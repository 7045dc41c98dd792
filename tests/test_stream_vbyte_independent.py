"""Stream VByte, stated a THIRD time and independently: a pure-Python codec written from the published format
(Lemire, Kurz, Rupp: "Stream VByte: Faster Byte-Oriented Integer Compression", the layout the `stream-vbyte` crate's
`Scalar` encoder documents) and hand-worked byte vectors, checked against BOTH the oracle (oracle/granne_oracle.cpp)
and the product's host code (granne_b200/csrc/formats.hpp).  The oracle and formats.hpp were written by the same hand;
this file shares no code with either, so a common misreading of the byte layout would have to be made three times.

Reference call sites: src/slice_vector/set_vector.rs:101 (`decode::<Scalar>`), :134 (`encode::<Scalar>`); crate
stream-vbyte 0.3.2 (Cargo.toml:42, not vendored; no rustc in this image, so these bytes are NOT pinned against a run
of the crate itself — stated in DESIGN.md §2).

Format: for n numbers, ceil(n/4) control bytes come first, then the data bytes.  Number i has a 2-bit code
(byte length - 1) in control byte i/4 at bit position 2*(i%4) (the first number of a quad in the LOW bits); its value
follows in the data stream as 1..4 little-endian bytes; 0 is encoded in one byte.
"""
import json

import numpy as np
import pytest

import granne_b200


# ---- the independent statement ---------------------------------------------------------------------------------------
def svb_len(v):
    return 1 if v < (1 << 8) else 2 if v < (1 << 16) else 3 if v < (1 << 24) else 4


def svb_encode(nums):
    ctrl = bytearray((len(nums) + 3) // 4)
    data = bytearray()
    for i, v in enumerate(nums):
        nb = svb_len(v)
        ctrl[i // 4] |= (nb - 1) << (2 * (i % 4))
        data += int(v).to_bytes(4, "little")[:nb]
    return bytes(ctrl) + bytes(data)


def svb_decode(buf, n):
    nctrl = (n + 3) // 4
    pos = nctrl
    out = []
    for i in range(n):
        nb = ((buf[i // 4] >> (2 * (i % 4))) & 3) + 1
        out.append(int.from_bytes(buf[pos:pos + nb], "little"))
        pos += nb
    return out, pos


def set_encode_independent(sorted_ids):
    """set_encode (set_vector.rs:117-148): count byte; deltas (first id as is); Stream VByte of max(4, count) numbers
    (zero padded) unless that is not smaller than 4*count raw little-endian bytes."""
    ids = list(sorted_ids)
    deltas = [ids[0]] + [b - a for a, b in zip(ids, ids[1:])] if ids else []
    padded = deltas + [0] * (4 - len(deltas)) if len(deltas) < 4 else deltas
    enc = svb_encode(padded)
    body = enc if len(enc) < 4 * len(ids) else b"".join(int(d).to_bytes(4, "little") for d in deltas)
    return bytes([len(ids)]) + body


def index_image_independent(layers):
    """Index::write_index (io.rs:11-70) + Offsets (offsets.rs:148-241) for explicit adjacency lists."""
    blobs = []
    for lists in layers:
        enc = [set_encode_independent(sorted(l)) for l in lists]
        offsets = [0]
        for e in enc:
            offsets.append(offsets[-1] + len(e))
        chunks = bytearray()
        for c in range(1 + len(lists) // 60):
            offs = offsets[c * 60:(c + 1) * 60]
            initial = offs[0] if offs else 0
            chunks += int(initial).to_bytes(8, "little")
            prev = initial
            for i in range(60):
                if i < len(offs):
                    chunks += int(offs[i] - prev).to_bytes(2, "little")
                    prev = offs[i]
                else:
                    chunks += b"\xff\xff"
        blobs.append(len(chunks).to_bytes(8, "little") + bytes(chunks) + b"".join(enc))
    meta = "granne" + json.dumps({"compressed": True, "granne_version": "0.5.2",
                                  "layer_counts": [len(l) for l in layers], "layer_sizes": [len(b) for b in blobs],
                                  "num_elements": len(layers[-1]) if layers else 0, "num_layers": len(layers),
                                  "num_neighbors": len(layers[-1][0]) if layers else 0, "version": 2},
                                 separators=(",", ":"))
    return meta.encode().ljust(1024, b" ") + b"".join(blobs)


# ---- hand-worked vectors ---------------------------------------------------------------------------------------------
HAND = [
    # numbers, bytes worked out by hand from the format description
    ([1, 2, 3, 4], bytes([0b00000000, 1, 2, 3, 4])),
    ([1, 300, 70000, (1 << 24) + 5],
     bytes([0b11100100, 0x01, 0x2C, 0x01, 0x70, 0x11, 0x01, 0x05, 0x00, 0x00, 0x01])),
    ([0, 255, 256, 65535, 65536], bytes([0b01010000, 0b00000010, 0x00, 0xFF, 0x00, 0x01, 0xFF, 0xFF,
                                         0x00, 0x00, 0x01])),
    ([0xFFFFFFFF, 0, 0, 0], bytes([0b00000011, 0xFF, 0xFF, 0xFF, 0xFF, 0, 0, 0])),
]


@pytest.mark.parametrize("nums,expected", HAND)
def test_hand_worked_vectors(nums, expected):
    assert svb_encode(nums) == expected
    back, used = svb_decode(expected, len(nums))
    assert back == nums and used == len(expected)


def _random_lists(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        deg = int(rng.integers(0, 41))
        hi = int(rng.choice([50, 300, 70_000, 20_000_000, 4_000_000_000]))
        out.append(sorted(set(int(x) for x in rng.integers(0, hi, size=deg))))
    return out


def test_oracle_list_codec_matches_the_independent_codec(oracle):
    lists = _random_lists(1, 400) + [[], [0], [7], [0, 1, 2, 3], [5, 70_000], list(range(0, 255 * 3, 3)),
                                     [0xFFFFFFFE], [1 << 24, (1 << 24) + 1, (1 << 25)]]
    for l in lists:
        mine = set_encode_independent(l)
        assert oracle.set_encode(l) == mine, l
        assert oracle.set_decode(mine) == l
        # and the independent decoder reads the oracle's bytes
        enc = oracle.set_encode(l)
        count = enc[0]
        if len(enc) - 1 == 4 * count:
            deltas = [int.from_bytes(enc[1 + 4 * i:5 + 4 * i], "little") for i in range(count)]
        else:
            deltas, _ = svb_decode(enc[1:], max(4, count))
            deltas = deltas[:count]
        assert list(np.cumsum(deltas, dtype=np.uint64)) == l


def test_product_reader_and_writer_match_the_independent_image():
    """formats.hpp (through the host-only C-ABI entry points, no GPU needed): the loader decodes an image written by
    the independent writer, and the library's own writer reproduces that image byte for byte."""
    granne_b200.load_library()
    rng = np.random.default_rng(5)
    for n_nodes in (1, 59, 60, 61, 240, 1000):
        lists = []
        for i in range(n_nodes):
            deg = int(rng.integers(1, 31))
            lists.append(sorted(set(int(x) for x in rng.integers(0, n_nodes, size=deg))) or [0])
        upper = lists[:max(1, n_nodes // 15)]
        upper = [[x for x in l if x < len(upper)] or [0] for l in upper]
        image = index_image_independent([upper, lists])
        shape = granne_b200.api.inspect_index(image)
        assert [s[0] for s in shape] == [len(upper), n_nodes]
        for layer, ref in enumerate([upper, lists]):
            rows = granne_b200.api.decode_layer(image, layer)
            for i, l in enumerate(ref):
                got = [int(x) for x in rows[i] if x != 0xFFFFFFFF]
                assert got == l, (n_nodes, layer, i)
        again = granne_b200.api.reencode_index(image)
        # the header's num_neighbors is "degree of node 0 in the last layer" in both writers; everything else is bytes
        assert again[1024:] == image[1024:]
        assert json.loads(again[6:1024].decode()) == json.loads(image[6:1024].decode())
