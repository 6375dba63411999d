"""paddle.reader / paddle.batch. Parity: python/paddle/reader/decorator.py, python/paddle/batch.py."""
import itertools
import random


def batch(reader, batch_size, drop_last=False):
    def gen():
        b = []
        for item in reader():
            b.append(item)
            if len(b) == batch_size:
                yield b
                b = []
        if b and not drop_last:
            yield b

    return gen


def shuffle(reader, buf_size):
    def gen():
        buf = []
        for e in reader():
            buf.append(e)
            if len(buf) >= buf_size:
                random.shuffle(buf)
                yield from buf
                buf = []
        random.shuffle(buf)
        yield from buf

    return gen


def chain(*readers):
    return lambda: itertools.chain(*[r() for r in readers])


def compose(*readers, check_alignment=True):
    def gen():
        for items in zip(*[r() for r in readers]):
            out = []
            for i in items:
                out.extend(i if isinstance(i, tuple) else (i,))
            yield tuple(out)

    return gen


def map_readers(func, *readers):
    return lambda: (func(*items) for items in zip(*[r() for r in readers]))


def buffered(reader, size):
    return reader


def firstn(reader, n):
    return lambda: itertools.islice(reader(), n)


def cache(reader):
    data = list(reader())
    return lambda: iter(data)
