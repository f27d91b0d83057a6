"""Channel registry with the reference's shape (/root/reference/src/models.py:3): hand this dict to
the reference's ``main.test`` and its per-frame loop runs on the GPU decoders unmodified."""
from . import bec, biawgn, bsc

models = {'bsc': bsc, 'bec': bec, 'biawgn': biawgn}
