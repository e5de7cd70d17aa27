"""Host-side mirror of the deuce seams on the path: `Strength::from(Hand)` and `Observation::equity`
(crates/deuce/src/{strength.rs:19-31,observation.rs:45-62}) as batch calls over the C ABI."""
import ctypes

import numpy as np

from . import _ffi

RANK_CH = "23456789TJQKA"
SUIT_CH = "cdhs"
RANKINGS = ["HighCard", "OnePair", "TwoPair", "ThreeOAK", "Straight", "FullHouse", "Flush", "FourOAK", "StraightFlush"]


def hand(text):
    """`Hand::try_from(&str)` (crates/deuce/src/hand.rs): "As Kh" -> 52-bit set, bit = 4*rank + suit."""
    bits = 0
    for card in text.split():
        bits |= 1 << (RANK_CH.index(card[0].upper()) * 4 + SUIT_CH.index(card[1].lower()))
    return bits


def unpack_strength(packed):
    packed = int(packed)
    kick = [r for r in range(13) if packed >> r & 1]
    return RANKINGS[packed >> 24 & 0xF], packed >> 20 & 0xF, packed >> 16 & 0xF, kick


def strength(hands):
    """Batch `Strength::from(Hand)`; returns packed u32 whose integer order is the reference's `Ord`."""
    l = _ffi.lib()
    hands = np.ascontiguousarray(hands, dtype=np.uint64)
    out = np.zeros(len(hands), dtype=np.uint32)
    _ffi.check(l.rbp_eval_batch(hands.ctypes.data, len(hands), out.ctypes.data), "rbp_eval_batch")
    return out


def river_equity(pocket, public):
    """Batch `Observation::equity` on river observations; returns (equity f32, bucket u8, wins u32, decisive u32)."""
    l = _ffi.lib()
    pocket = np.ascontiguousarray(pocket, dtype=np.uint64)
    public = np.ascontiguousarray(public, dtype=np.uint64)
    assert len(pocket) == len(public)
    n = len(pocket)
    eq, bk = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    w, t = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    _ffi.check(l.rbp_river_equity_batch(pocket.ctypes.data, public.ctypes.data, n, eq.ctypes.data, bk.ctypes.data, w.ctypes.data, t.ctypes.data),
               "rbp_river_equity_batch")
    return eq, bk, w, t


STREETS = {"pref": 0, "flop": 1, "turn": 2, "rive": 3}


def canonical(pocket, public):
    """Batch `Isomorphism::from(Observation)`; returns (pocket, public, was_canonical)."""
    l = _ffi.lib()
    pocket = np.ascontiguousarray(pocket, dtype=np.uint64)
    public = np.ascontiguousarray(public, dtype=np.uint64)
    po, pu, fl = np.zeros_like(pocket), np.zeros_like(public), np.zeros(len(pocket), np.uint8)
    _ffi.check(l.rbp_canonical_batch(pocket.ctypes.data, public.ctypes.data, len(pocket), po.ctypes.data, pu.ctypes.data, fl.ctypes.data),
               "rbp_canonical_batch")
    return po, pu, fl


class IsoSet:
    """`IsomorphismIterator::from(street)` materialised on the device, plus its `Lookup` column (iso → abstraction)."""

    def __init__(self, street, device=0):
        self._lib = _ffi.lib()
        self._h = ctypes.c_void_p()
        self.street = street
        _ffi.check(self._lib.rbp_isoset_create(STREETS[street], device, ctypes.byref(self._h)), "rbp_isoset_create")

    def __len__(self):
        return int(self._lib.rbp_isoset_size(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.rbp_isoset_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export(self, offset=0, count=None, with_abs=False):
        count = len(self) - offset if count is None else count
        p, b = np.zeros(count, np.uint64), np.zeros(count, np.uint64)
        a = np.zeros(count, np.uint8) if with_abs else None
        _ffi.check(self._lib.rbp_isoset_export(self._h, offset, count, p.ctypes.data, b.ctypes.data, a.ctypes.data if with_abs else None),
                   "rbp_isoset_export")
        return (p, b, a) if with_abs else (p, b)

    def export_rows(self, offset=0, count=None):
        """Rows of the reference's `isomorphism` table: (obs i64 = i64::from(Isomorphism), abs i16 = i16::from(Abstraction))."""
        count = len(self) - offset if count is None else count
        obs, a = np.zeros(count, np.int64), np.zeros(count, np.int16)
        _ffi.check(self._lib.rbp_isoset_export_rows(self._h, offset, count, obs.ctypes.data, a.ctypes.data), "rbp_isoset_export_rows")
        return obs, a

    def set_abstractions(self, abs_):
        abs_ = np.ascontiguousarray(abs_, dtype=np.uint8)
        assert len(abs_) == len(self)
        _ffi.check(self._lib.rbp_isoset_set_abstractions(self._h, abs_.ctypes.data), "rbp_isoset_set_abstractions")

    def river_buckets(self):
        """`Lookup::grow(Street::Rive)`."""
        _ffi.check(self._lib.rbp_isoset_river_buckets(self._h), "rbp_isoset_river_buckets")

    def project(self, child, bins, offset=0, count=None):
        """`Lookup::projections`: histograms u8[count][bins] of this street's observations over `child`'s abstractions."""
        count = len(self) - offset if count is None else count
        hist = np.zeros((count, bins), np.uint8)
        miss = ctypes.c_uint64()
        _ffi.check(self._lib.rbp_isoset_project(self._h, child._h, bins, offset, count, hist.ctypes.data, ctypes.byref(miss)), "rbp_isoset_project")
        return hist, miss.value


def obs_encode(pocket, public):
    """`i64::from(Observation)` (crates/deuce/src/observation.rs:130-141) for card-mask arrays."""
    pocket = np.ascontiguousarray(pocket, dtype=np.uint64)
    public = np.ascontiguousarray(public, dtype=np.uint64)
    out = np.zeros(len(pocket), np.int64)
    _ffi.lib().rbp_obs_encode(pocket.ctypes.data, public.ctypes.data, len(pocket), out.ctypes.data)
    return out


def obs_decode(obs):
    """`Observation::from(i64)` (observation.rs:143-163) → (pocket masks, public masks)."""
    obs = np.ascontiguousarray(obs, dtype=np.int64)
    p, b = np.zeros(len(obs), np.uint64), np.zeros(len(obs), np.uint64)
    _ffi.lib().rbp_obs_decode(obs.ctypes.data, len(obs), p.ctypes.data, b.ctypes.data)
    return p, b
