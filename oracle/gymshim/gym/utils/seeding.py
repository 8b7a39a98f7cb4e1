"""gym 0.21 seeding algorithm: sha512-hashed seed -> legacy numpy RandomState."""
import hashlib
import os
import struct

import numpy as np


def _bigint_from_bytes(data):
    sizeof_int = 4
    padding = sizeof_int - len(data) % sizeof_int
    data += b'\0' * padding
    int_count = len(data) // sizeof_int
    unpacked = struct.unpack(f'{int_count}I', data)
    return sum(2 ** (sizeof_int * 8 * i) * val for i, val in enumerate(unpacked))


def create_seed(a=None, max_bytes=8):
    if a is None:
        return _bigint_from_bytes(os.urandom(max_bytes))
    if isinstance(a, int):
        return a % 2 ** (8 * max_bytes)
    raise TypeError(f'Invalid type for seed: {type(a)} ({a})')


def hash_seed(seed=None, max_bytes=8):
    if seed is None:
        seed = create_seed(max_bytes=max_bytes)
    digest = hashlib.sha512(str(seed).encode('utf8')).digest()
    return _bigint_from_bytes(digest[:max_bytes])


def _int_list_from_bigint(bigint):
    if bigint < 0:
        raise ValueError(f'Seed must be non-negative, not {bigint}')
    if bigint == 0:
        return [0]
    ints = []
    while bigint > 0:
        bigint, mod = divmod(bigint, 2 ** 32)
        ints.append(mod)
    return ints


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, (int, np.integer)) and seed >= 0):
        raise ValueError(f'Seed must be a non-negative integer or omitted, not {seed}')
    seed = create_seed(None if seed is None else int(seed))
    rng = np.random.RandomState()
    rng.seed(_int_list_from_bigint(hash_seed(seed)))
    return rng, seed
