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
