"""Host-side mirror of ORBVocabulary (include/ORBVocabulary.h:31-32 = DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB>)
for the one operation on the hot path: transform(features, BowVector, FeatureVector, levelsup) as called by
Frame::ComputeBoW (Frame.cc:425-432) / KeyFrame::ComputeBoW (KeyFrame.cc:59-70).  The descent, the grouping and the
normalisation run in liborbgpu.so (csrc/og_vocab.cu); nothing here computes — no CUDA device, no result."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

_bound = False


def _lib():
    global _bound
    L = capi.lib()
    if not _bound:
        vp, i = C.c_void_p, C.c_int
        L.orbgpu_vocabulary_create.argtypes = [C.POINTER(vp), i, i, i, i, i, i, vp, vp, vp, vp]
        L.orbgpu_vocabulary_destroy.argtypes = [vp]
        L.orbgpu_vocabulary_info.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
        L.orbgpu_vocabulary_sync.argtypes = [vp]
        L.orbgpu_vocabulary_last_launches.argtypes = [vp]
        L.orbgpu_vocabulary_stream.argtypes = [vp, C.POINTER(vp)]
        L.orbgpu_bow_transform.argtypes = [vp, i, vp, vp, i] + [vp] * 9
        L.orbgpu_bow_transform_dev.argtypes = [vp, i, vp, i, i, vp, i] + [vp] * 9
        _bound = True
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def read_text_file(path: str):
    """The records of a vocabulary text file (format of TemplatedVocabulary::loadFromTextFile / saveToTextFile,
    TemplatedVocabulary.h:1338-1450): header `k L scoring weighting`, then `parent isLeaf d0 .. d31 weight` per node."""
    with open(path) as fh:
        head = fh.readline().split()
        k, L, scoring, weighting = (int(x) for x in head[:4])
        if k < 0 or k > 20 or L < 1 or L > 10 or scoring < 0 or scoring > 5 or weighting < 0 or weighting > 3:
            raise ValueError("Vocabulary loading failure: This is not a correct text file!")   # :1359-1363
        rows = [ln.split() for ln in fh if ln.strip()]
    parent = np.array([int(r[0]) for r in rows], np.int32)
    is_leaf = np.array([1 if int(r[1]) > 0 else 0 for r in rows], np.uint8)
    desc = np.array([[int(x) for x in r[2:34]] for r in rows], np.uint8).reshape(-1, 32)
    weight = np.array([float(r[34]) for r in rows], np.float64)
    return {"k": k, "L": L, "scoring": scoring, "weighting": weighting, "parent": parent, "is_leaf": is_leaf, "desc": desc, "weight": weight}


def write_text_file(path: str, voc, scoring: int = 0, weighting: int = 0):
    """saveToTextFile (TemplatedVocabulary.h:1428-1450); weights with 17 significant digits so they read back exactly.
    No newline after the last record: the reference's loader loops `while(!f.eof())` (:1378) and would read an empty last
    line into uninitialised `pid` / `nIsLeaf` (:1389-1395).  read_text_file skips blank lines instead."""
    with open(path, "w") as fh:
        fh.write(f"{voc['k']} {voc['L']}  {scoring} {weighting}")
        for p, l, d, w in zip(voc["parent"], voc["is_leaf"], voc["desc"], voc["weight"]):
            fh.write(f"\n{int(p)} {int(l)} " + " ".join(str(int(x)) for x in d) + f"  {float(w)!r}")


class ORBVocabulary:
    """`ORBVocabulary voc; voc.loadFromTextFile(path); voc.transform(descriptors, levelsup=4)`."""

    def __init__(self, device: int = 0):
        self.device = device
        self._h = C.c_void_p(None)
        self.k = self.L = 0

    def loadFromTextFile(self, path: str) -> bool:
        rec = read_text_file(path)
        self.from_records(rec, rec["scoring"], rec["weighting"])
        return True

    def from_records(self, rec, scoring: int = 0, weighting: int = 0):
        self.close()
        parent = np.ascontiguousarray(rec["parent"], np.int32)
        is_leaf = np.ascontiguousarray(rec["is_leaf"], np.uint8)
        desc = np.ascontiguousarray(rec["desc"], np.uint8)
        weight = np.ascontiguousarray(rec["weight"], np.float64)
        h = C.c_void_p(None)
        capi.check(_lib().orbgpu_vocabulary_create(C.byref(h), self.device, int(rec["k"]), int(rec["L"]), scoring, weighting, len(parent),
                                                  _ptr(parent), _ptr(is_leaf), _ptr(desc), _ptr(weight)))
        self._h = h
        self.k, self.L = int(rec["k"]), int(rec["L"])
        return self

    def close(self):
        if self._h:
            _lib().orbgpu_vocabulary_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self) -> int:
        n, w = C.c_int(0), C.c_int(0)
        capi.check(_lib().orbgpu_vocabulary_info(self._h, C.byref(n), C.byref(w)))
        return w.value

    def empty(self) -> bool:
        return not self._h or self.size() == 0

    @property
    def last_launches(self) -> int:
        return _lib().orbgpu_vocabulary_last_launches(self._h)

    def transform_batch(self, kp_off, desc, levelsup: int = 4):
        """All frames of a batch in one call.  Returns a dict of the CSR arrays of include/orbgpu.h."""
        if not self._h:
            raise capi.OrbGpuError("vocabulary is empty (loadFromTextFile / from_records first)")
        kp_off = np.ascontiguousarray(kp_off, np.int32)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        nf, n = len(kp_off) - 1, int(kp_off[-1])
        out = {"bv_off": np.zeros(nf + 1, np.int32), "bv_word": np.zeros(n, np.uint32), "bv_value": np.zeros(n, np.float64),
               "fv_node_off": np.zeros(nf + 1, np.int32), "fv_node_id": np.zeros(n, np.int32), "fv_feat_off": np.zeros(n + 1, np.int32),
               "fv_feat": np.zeros(n, np.int32), "word_of_feature": np.zeros(n, np.uint32), "node_of_feature": np.zeros(n, np.uint32)}
        capi.check(_lib().orbgpu_bow_transform(self._h, nf, _ptr(kp_off), _ptr(desc), levelsup, *[_ptr(out[k]) for k in
                                               ("bv_off", "bv_word", "bv_value", "fv_node_off", "fv_node_id", "fv_feat_off", "fv_feat",
                                                "word_of_feature", "node_of_feature")]))
        nw, nn = int(out["bv_off"][-1]), int(out["fv_node_off"][-1])
        nv = int(out["fv_feat_off"][nn])
        out["bv_word"], out["bv_value"] = out["bv_word"][:nw], out["bv_value"][:nw]
        out["fv_node_id"], out["fv_feat_off"], out["fv_feat"] = out["fv_node_id"][:nn], out["fv_feat_off"][:nn + 1], out["fv_feat"][:nv]
        return out

    def transform(self, desc, levelsup: int = 4):
        """One frame: (BowVector as {word: value}, FeatureVector as {node: [feature indices]}) — the two std::maps of the reference."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        o = self.transform_batch(np.array([0, len(desc)], np.int32), desc, levelsup)
        bv = {int(w): float(v) for w, v in zip(o["bv_word"], o["bv_value"])}
        fo = o["fv_feat_off"]
        fv = {int(nid): [int(x) for x in o["fv_feat"][fo[j]:fo[j + 1]]] for j, nid in enumerate(o["fv_node_id"])}
        return bv, fv
