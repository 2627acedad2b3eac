"""CPU restatement of the reference batch loader (training/loader.py) — TEST INFRASTRUCTURE ONLY.

Plain Python / numpy following the reference line by line (list-of-ints sequences, list-of-lists batches); imported
only by tests/ and scripts/bench_loader.py as the checker / CPU baseline of omnibiote_b200/loader.py. Pinned against
batches produced by the unmodified reference (oracle/gen_loader_golden.py -> tests/golden/loader_*.npz).
"""
import numpy as np

EOS_TOKEN, MASK_TOKEN, PAD_TOKEN = 3, 2, 1  # loader.py:4-6


def line_reader(filenames, banned_tokens):
    """loader.py:25-58"""
    while True:
        np.random.shuffle(filenames)                                   # :31
        chunk_size = 10                                                # :33
        chunked = np.split(filenames, np.arange(chunk_size, len(filenames), chunk_size))  # :35
        for name in chunked:
            block = np.concatenate([np.load(f) for f in name])         # :38-42
            eos_indices = np.where(block == EOS_TOKEN)[0]              # :43
            sub_blocks = np.split(block, eos_indices + 1)              # :44
            order = np.arange(len(sub_blocks))                         # :47
            np.random.shuffle(order)                                   # :48
            for idx in order:
                sub_block = sub_blocks[idx]
                if len(sub_block) > 0:                                 # :52
                    if len(banned_tokens) == 1:
                        mask = sub_block != banned_tokens[0]           # :55
                    else:
                        mask = ~np.isin(sub_block, banned_tokens)      # :57
                    yield np.int32(sub_block[mask])                    # :58-59


def get_sequence(reader, ctx_len, USE_PADDING=False):
    """loader.py:116-159 (the active definition)"""
    sequence = []
    while True:
        line = next(reader)
        seq_len = len(sequence)
        if seq_len == ctx_len:                                         # :129
            yield sequence
            sequence = []
            continue
        if seq_len + len(line) > ctx_len:                              # :138
            if USE_PADDING:
                if seq_len == 0:
                    continue                                           # :141-143
                sequence.extend([PAD_TOKEN] * (ctx_len - seq_len))     # :146
            else:
                sequence.extend(line[:ctx_len - seq_len])              # :149
            yield sequence
            sequence = []
            continue
        sequence.extend(line)                                          # :157


def get_batch(generators, train_ints):
    """loader.py:161-181 with return_pt=False"""
    while True:
        batch = []
        for generator, train_int in zip(generators, train_ints):
            for _ in range(train_int):
                batch.append(next(generator))
        np.random.shuffle(batch)                                       # :173
        yield np.asarray(batch)
