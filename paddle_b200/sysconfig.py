"""paddle.sysconfig. Parity: python/paddle/sysconfig.py."""
import os


def get_include():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "include")


def get_lib():
    return os.path.dirname(os.path.abspath(__file__))
