"""Batch loader of the pre-training step: a vectorised drop-in for the reference's ``training/loader.py``
(``line_reader`` / ``get_sequence`` / ``get_batch`` / ``data_loader_parallel``, used at train_encoder.py:5,127-142).

Same generators, same argument meaning, same consumption of numpy's global RNG (``np.random.shuffle`` of the file list,
of the document order of every chunk and of every batch), so with the same seed and shards the batches are identical
to the reference's, element for element (tests/test_loader_cpu.py, golden batches from the unmodified reference).
What changes is the representation: the reference builds every sequence as a Python list of Python ints
(``sequence.extend(line)``, loader.py:139-159) and every batch with ``torch.tensor(list_of_lists)`` (loader.py:179);
here a sequence is one preallocated int32 row filled by slice copies and a batch is one contiguous array handed to
torch without a per-element conversion (optionally in pinned memory for an asynchronous H2D copy).

Reference quirks kept on purpose (they decide which tokens reach the model):
  * when a sequence is exactly full, the line read next is dropped (loader.py:129-135);
  * without padding the tail of a truncated line is dropped, with padding the line that did not fit is dropped, and a
    line longer than ctx_len arriving on an empty sequence is skipped (loader.py:138-153);
  * the last piece of a chunk (tokens after its final EOS) is a document of its own (np.split, loader.py:44).
"""
from __future__ import annotations

import numpy as np
import torch

EOS_TOKEN = 3
MASK_TOKEN = 2
PAD_TOKEN = 1

CHUNK_FILES = 10  # loader.py:34: files loaded and shuffled together


def data_loader_parallel(batch_queue, batch_generator, device):
    """Thread body of train_encoder.py:140-142 (loader.py:8-23): move batches to ``device`` and queue them."""
    while True:
        try:
            data = next(batch_generator)
            data = data.to(device, non_blocking=True)
            batch_queue.put(data)
        except StopIteration:
            break


def line_reader(filenames, banned_tokens):
    """Yields one document (np.int32 array, EOS included) at a time, forever (loader.py:25-58)."""
    banned = np.asarray(list(banned_tokens))
    while True:
        np.random.shuffle(filenames)
        chunks = np.split(filenames, np.arange(CHUNK_FILES, len(filenames), CHUNK_FILES))
        for names in chunks:
            block = np.concatenate([np.load(f) for f in names])
            eos = np.flatnonzero(block == EOS_TOKEN)
            starts = np.concatenate(([0], eos + 1))
            ends = np.concatenate((eos + 1, [len(block)]))
            order = np.arange(len(starts))
            np.random.shuffle(order)
            if len(banned) == 1:
                bad = block == banned[0]
            elif len(banned) > 1:
                bad = np.isin(block, banned)
            else:
                bad = None
            any_bad = bad is not None and bool(bad.any())
            block32 = block.astype(np.int32, copy=False)
            for idx in order:
                s, e = starts[idx], ends[idx]
                if e > s:
                    doc = block32[s:e]
                    if any_bad:
                        doc = doc[~bad[s:e]]
                    yield doc


def get_sequence(reader, ctx_len, USE_PADDING=False):
    """Packs documents into rows of exactly ``ctx_len`` tokens (loader.py:116-159). Yields np.int32 arrays."""
    row = np.empty(ctx_len, dtype=np.int32)
    n = 0
    while True:
        line = next(reader)
        if n == ctx_len:  # full: hand it out; the line just read is dropped, as in the reference
            yield row
            row = np.empty(ctx_len, dtype=np.int32)
            n = 0
            continue
        m = len(line)
        if n + m > ctx_len:
            if USE_PADDING:
                if n == 0:
                    continue  # a line longer than ctx_len on an empty row is skipped
                row[n:] = PAD_TOKEN
            else:
                row[n:] = line[:ctx_len - n]
            yield row
            row = np.empty(ctx_len, dtype=np.int32)
            n = 0
            continue
        row[n:n + m] = line
        n += m


def get_batch(generators, train_ints, return_pt=False, device="cpu", pinned=False):
    """``train_ints[k]`` rows from ``generators[k]``, shuffled (loader.py:161-181). ``return_pt``: int64 torch tensor
    on ``device`` like the reference; ``pinned`` keeps the host copy in page-locked memory (asynchronous H2D)."""
    total = int(sum(train_ints))
    while True:
        rows = []
        for generator, train_int in zip(generators, train_ints):
            for _ in range(train_int):
                rows.append(next(generator))
        batch = np.stack(rows) if rows else np.empty((0, 0), dtype=np.int32)
        assert batch.shape[0] == total
        np.random.shuffle(batch)  # same draws as shuffling the reference's list of rows
        if return_pt:
            t = torch.from_numpy(batch.astype(np.int64))
            if pinned and torch.cuda.is_available():
                t = t.pin_memory()
            yield t if str(device) == "cpu" else t.to(device, non_blocking=pinned)
        else:
            yield batch
