#!/usr/bin/env python
"""Longer fuzz of the host parsers (sparse buffers from the wire, band bit streams, random prefix-free code sets with long
code words) than the unit tests run.  Meant for an instrumented build: `tools/sanitize_host.sh --fuzz N` swaps in the
ASan + UBSan library and runs this with N rounds.  No GPU needed."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def random_code_set(rng, n_values, n_runs, max_len):
    """A random prefix-free set: leaves of a random binary tree, longest codes up to max_len bits."""
    leaves, frontier = [], [""]
    want = n_values + n_runs + 1
    while len(leaves) + len(frontier) < want or (frontier and rng.random() < 0.3):
        if not frontier:
            break
        i = int(rng.integers(0, len(frontier)))
        code = frontier.pop(i)
        if len(code) >= max_len - 1:
            leaves.append(code + "0"); leaves.append(code + "1")
        else:
            frontier += [code + "0", code + "1"]
        if len(leaves) + len(frontier) > 4 * want:
            break
    leaves += frontier
    rng.shuffle(leaves)
    leaves = [c for c in leaves if 1 <= len(c) <= 31][:want]
    kinds = [2] + [1] * min(n_runs, len(leaves) - 2) + [0] * max(len(leaves) - 1 - n_runs, 1)
    kinds = kinds[:len(leaves)]
    args = [0 if k == 2 else int(rng.integers(1, 300)) if k == 1 else int(rng.integers(-2000, 2000)) for k in kinds]
    return leaves, kinds, args


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    pkg = importlib.import_module("cineform-sdk_b200")
    rng = np.random.default_rng(2024)
    lay = pkg.layout_for(pkg.FrameDesc(704, 96, pkg.PIXEL_YUYV))
    words = lay.coded_bytes // 2
    accepted = rejected = 0
    for r in range(rounds):
        leaves, kinds, args = random_code_set(rng, int(rng.integers(4, 200)), int(rng.integers(1, 40)), int(rng.integers(4, 30)))
        try:
            book = pkg.VlcDecodebook.from_arrays([int(c, 2) for c in leaves], [len(c) for c in leaves], kinds, args)
            dec = pkg.VlcDecoder(lay, book)
        except pkg.CfbError:
            rejected += 1
            continue
        out = np.zeros(pkg.sparse_max_bytes(lay), np.uint8)
        for trial in range(20):
            dec.begin(out)
            # half of the streams are built from valid code words (so that long sequences are parsed), half are noise
            if trial % 2:
                picks = rng.integers(0, len(leaves), int(rng.integers(1, 3000)))
                text = "".join(leaves[i] for i in picks if kinds[i] != 2) + (leaves[kinds.index(2)] if rng.random() < 0.7 else "")
                text += "0" * (-len(text) % 8)
                stream = np.array([int(text[i:i + 8], 2) for i in range(0, len(text), 8)] or [0], np.uint8)
                if rng.random() < 0.3 and stream.size > 2:
                    stream = stream[:int(rng.integers(1, stream.size))]
            else:
                stream = rng.integers(0, 256, int(rng.integers(1, 600))).astype(np.uint8)
            try:
                c = int(rng.integers(0, 3)); k = int(rng.integers(0, 3)); b = int(rng.integers(1, 4))
                dec.band(c, k, b, stream, int(rng.integers(1, 60)))
                back = pkg.sparse_expand(lay, dec.end())
                assert back.size == lay.coded_bytes
                accepted += 1
            except pkg.CfbError:
                rejected += 1
        dec.close()
        # damaged sparse buffers
        dense = np.where(rng.random(words) < rng.random() * 0.3, rng.integers(-3000, 3000, words), 0).astype(np.int16)
        good = pkg.sparse_compact(lay, dense.view(np.uint8))
        for trial in range(10):
            bad = np.zeros(pkg.sparse_max_bytes(lay), np.uint8)
            bad[:good.size] = good
            n = int(rng.integers(1, 8))
            bad[rng.integers(0, good.size, n)] = rng.integers(0, 256, n).astype(np.uint8)
            for call in (lambda: pkg.sparse_expand(lay, bad), lambda: pkg.sparse_band_nonzeros(lay, bad, int(rng.integers(0, 3)), int(rng.integers(0, 3)), 1),
                         lambda: pkg.sparse_expand_band(lay, bad, 1, 2, 0)):
                try:
                    call()
                    accepted += 1
                except pkg.CfbError:
                    rejected += 1
    print(f"fuzz: {rounds} rounds, {accepted} inputs accepted, {rejected} rejected, no crash")


if __name__ == "__main__":
    main()
