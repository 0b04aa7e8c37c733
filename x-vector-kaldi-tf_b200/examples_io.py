"""Host-side mirror of the training-example readers of the reference's ``local/tf/examples_io.py``.

Only what feeds ``Model.train_one_iteration`` / ``Model.eval`` is here:

  * ``TarFileDataLoader`` (reference examples_io.py:224-255): an ``egs.N.tar`` whose members
    ``minibatch_<i>.npy`` are float16 ``[minibatch, len, D]`` arrays (written at examples_io.py:159-178) and the
    sibling ``egs.N.npy`` holding one int label vector per minibatch (create_tar_files.py:130-134; an object array,
    hence ``allow_pickle``).  A daemon thread decodes members into a bounded queue; ``pop()`` returns
    ``(float array [B, T, D], labels [B])`` and ``count`` is the number of minibatches.
  * ``ArrayDataLoader``: the same ``count`` / ``pop()`` contract over in-memory minibatches (the reference's
    ``DataLoader`` with ``sequential_loading=False``, examples_io.py:181-222, minus the scp/ranges plumbing).
  * ``write_egs_tar``: writes that on-disk format (tests and synthetic benchmarks; mirrors ``save_data_info_tar``).
"""
from __future__ import annotations

import io
import queue
import tarfile
import time
from threading import Thread

import numpy as np


def write_egs_tar(tar_file_path, minibatches, labels):
    """minibatches: list of [B_i, T_i, D] arrays (stored as float16); labels: list of int vectors [B_i]."""
    assert tar_file_path.endswith(".tar") and len(minibatches) == len(labels)
    with tarfile.TarFile(tar_file_path, "w") as tar:
        for i, mat in enumerate(minibatches):
            buf = io.BytesIO()
            np.save(buf, np.asarray(mat).astype(np.float16))
            size = buf.tell()
            buf.seek(0)
            info = tarfile.TarInfo(name="minibatch_%d.npy" % i)
            info.size = size
            tar.addfile(tarinfo=info, fileobj=buf)
    lab = np.empty(len(labels), dtype=object)
    for i, v in enumerate(labels):
        lab[i] = np.asarray(v, dtype=np.int32)
    np.save(tar_file_path.replace(".tar", ".npy"), lab, allow_pickle=True)


class TarFileDataLoader(object):

    def __init__(self, tar_file, logger=None, queue_size=5):
        self._train_labels = np.load(tar_file.replace(".tar", ".npy"), allow_pickle=True)
        self._tar = tarfile.open(tar_file, "r")
        self._names = self._tar.getnames()
        self._total_count = len(self._names)
        self.count = self._total_count
        self._read_index = 0
        assert self._total_count == self._train_labels.shape[0]
        self._logger = logger
        self.queue = queue.Queue(queue_size)
        self._thread = Thread(target=self.__load_data)
        self._thread.daemon = True
        self._thread.start()

    def __load_data(self):
        try:
            while self._read_index < len(self._names):
                name = self._names[self._read_index]
                idx = int(name[:-4].split("_")[1])
                label = self._train_labels[idx]
                start_time = time.time()
                # the member is read whole, then parsed (np.load cannot take tarfile's file object on every numpy)
                mat = np.load(io.BytesIO(self._tar.extractfile(name).read()))
                if self._logger is not None:
                    self._logger.info("Loading one minibatch take %d seconds." % (time.time() - start_time))
                self.queue.put((mat, label))
                self._read_index += 1
        except BaseException as e:                 # surface reader failures in the consumer instead of a silent timeout
            self.queue.put(e)

    def pop(self, timeout=30):
        if self._total_count == 0:
            return None, None
        item = self.queue.get(block=True, timeout=timeout)
        if isinstance(item, BaseException):
            raise item
        return item


class ArrayDataLoader(object):
    """``count`` / ``pop()`` over minibatches already in memory (popped from the END, as the reference's list.pop())."""

    def __init__(self, minibatches, labels):
        assert len(minibatches) == len(labels)
        self.train_data = list(minibatches)
        self.train_labels = list(labels)
        self.count = len(self.train_data)

    def pop(self, timeout=30):
        if len(self.train_data) == 0:
            return None, None
        return self.train_data.pop(), self.train_labels.pop()
