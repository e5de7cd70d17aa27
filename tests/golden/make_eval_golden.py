"""Generates tests/golden/eval_known_answers.json from the reference's own evaluator tests
(crates/deuce/src/evaluator.rs:186-357).  Run in the build container (needs /root/reference); the JSON is committed
so the GPU box never reads the reference."""
import json
import os
import re

SRC = "/root/reference/crates/deuce/src/evaluator.rs"
RANKS = {"Two": 0, "Three": 1, "Four": 2, "Five": 3, "Six": 4, "Seven": 5, "Eight": 6, "Nine": 7, "Ten": 8, "Jack": 9, "Queen": 10,
         "King": 11, "Ace": 12}
RANK_CH = "23456789TJQKA"
SUIT_CH = "cdhs"
TAGS = {"HighCard": 0, "OnePair": 1, "TwoPair": 2, "ThreeOAK": 3, "Straight": 4, "FullHouse": 5, "Flush": 6, "FourOAK": 7, "StraightFlush": 8}


def hand_bits(text):
    bits = 0
    for card in text.split():
        bits |= 1 << (RANK_CH.index(card[0]) * 4 + SUIT_CH.index(card[1]))
    return bits


def main():
    text = open(SRC).read()
    # default (non-shortdeck) build only: the `#[cfg(not(feature = "shortdeck"))] mod tests` block
    start = text.index('#[cfg(not(feature = "shortdeck"))]\nmod tests')
    end = text.index('#[cfg(test)]', start + 10)
    text = text[start:end]
    cases = []
    for m in re.finditer(r'fn (\w+)\(\) \{\s*let eval = Evaluator::from\(Hand::try_from\("([^"]+)"\).*?'
                         r'assert_eq!\(ranking, Ranking::(\w+)\(([^)]*)\)\);\s*assert_eq!\(kickers, Kickers::from\(vec!\[([^\]]*)\]\)\);',
                         text, flags=re.S):
        name, hand, tag, ranks, kicks = m.groups()
        rs = [RANKS[r.strip().split("::")[1]] for r in ranks.split(",") if r.strip()]
        ks = [RANKS[k.strip().split("::")[1]] for k in kicks.split(",") if k.strip()]
        kick_bits = 0
        for k in ks:
            kick_bits |= 1 << k
        packed = TAGS[tag] << 24 | rs[0] << 20 | (rs[1] if len(rs) > 1 else 0) << 16 | kick_bits
        cases.append({"name": name, "hand": hand, "bits": hand_bits(hand), "ranking": tag, "ranks": rs, "kickers": ks, "packed": packed})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eval_known_answers.json")
    json.dump({"source": "crates/deuce/src/evaluator.rs:186-357", "cases": cases}, open(out, "w"), indent=1)
    print(len(cases), "cases ->", out)


if __name__ == "__main__":
    main()
