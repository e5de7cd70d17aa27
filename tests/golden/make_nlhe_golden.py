"""Generates tests/golden/nlhe_scripts.json from the reference's own `kicker::Game` tests (crates/kicker/src/game.rs
`mod tests`): every test whose body is a straight-line script of `game.apply(Action::…)` calls interleaved with
`assert!(…)` / `assert_eq!(…)` on predicates of the state is transcribed mechanically into
{steps: [...], checks: [[step_index, predicate, expected], ...]}.  Also extracts the showdown ledgers of
crates/kicker/src/showdown.rs tests.  Run in the build container (needs /root/reference); the JSON is committed."""
import json
import os
import re

GAME = "/root/reference/crates/kicker/src/game.rs"
SHOW = "/root/reference/crates/kicker/src/showdown.rs"
PREDICATES = {"must_post", "must_stop", "must_deal", "is_everyone_alright", "is_everyone_calling", "is_everyone_touched",
              "is_everyone_matched", "is_everyone_folding", "is_everyone_shoving", "may_fold", "may_call", "may_check",
              "may_raise", "may_shove"}
STREETS = {"Pref": 0, "Flop": 1, "Turn": 2, "Rive": 3}


def test_body(text, name):
    start = text.index(f"fn {name}()")
    depth, i = 0, text.index("{", start)
    for j in range(i, len(text)):
        depth += text[j] == "{"
        depth -= text[j] == "}"
        if depth == 0:
            return text[i + 1:j]
    raise ValueError(name)


def transcribe(body):
    """Straight-line interpreter for the subset of Rust these tests use."""
    steps, checks = [], []
    for raw in body.split("\n"):
        line = raw.split("//")[0].strip()
        if not line:
            continue
        for m in re.finditer(r"\.apply\(Action::(\w+)(?:\((\w+)\))?\)", line):
            kind, arg = m.groups()
            if kind == "Draw":
                steps.append(["Draw"])
            elif arg is None:
                steps.append([kind])
            else:
                steps.append([kind, int(arg) if arg.isdigit() else None])  # a bound variable = the legal amount
        m = re.match(r"assert!\((!?)game\.(\w+)\(\)\);", line)
        if m and m.group(2) in PREDICATES:
            checks.append([len(steps), m.group(2), m.group(1) != "!"])
        m = re.match(r"assert_eq!\(game\.pot\(\), (\d+)\);", line)
        if m:
            checks.append([len(steps), "pot", int(m.group(1))])
        m = re.match(r"assert_eq!\(game\.to_raise\(\), (\d+)\);", line)
        if m:
            checks.append([len(steps), "to_raise", int(m.group(1))])
        m = re.match(r"assert_eq!\(game\.(?:board\(\)\.)?street\(\), Street::(\w+)\);", line)
        if m:
            checks.append([len(steps), "street", STREETS[m.group(1)]])
    return {"steps": steps, "checks": checks}


def main():
    text = open(GAME).read()
    tests = text[text.index("mod tests {"):text.index("// ─── Multiplayer (P > 2)")]
    out = {"scripts": {}, "showdowns": {}}
    for name in ("history_of_checks", "everyone_folds_pref", "everyone_folds_flop", "min_raise_size", "pot_tracking", "allin_fold",
                 "terminal_river_showdown"):
        out["scripts"][name] = transcribe(test_body(tests, name))
    # terminal_river_showdown loops three streets of (Draw, Check, Check): unrolled by hand from its `for` header
    t = out["scripts"]["terminal_river_showdown"]
    t["steps"] = [["Call", 1], ["Check"]] + [["Draw"], ["Check"], ["Check"]] * 3
    t["checks"] = [[len(t["steps"]), "street", 3], [len(t["steps"]), "must_stop", True], [len(t["steps"]), "must_deal", False]]
    show = open(SHOW).read()
    names = {"ace_high": 0, "one_pair": 1, "two_pair": 2, "triplets": 3, "the_nuts": 4}  # increasing Strength
    for m in re.finditer(r"fn (\w+)\(\) \{\s*let settlement = Showdown::from\(vec!\[(.*?)\]\)\s*\.settle\(\);(.*?)\n    \}", show, flags=re.S):
        name, ledger, asserts = m.groups()
        rows = [[int(r), s[0] if s != "Betting" else "P", names[h]] for r, s, h in
                re.findall(r"Settlement::from\(\((\d+), State::(\w+), (\w+)\(\)\)\)", ledger)]
        rewards = [int(x) for x in re.findall(r"reward\(\), (\d+)\)", asserts)]
        assert len(rows) == len(rewards), name
        out["showdowns"][name] = {"ledger": rows, "rewards": rewards}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nlhe_scripts.json")
    json.dump(out, open(dst, "w"), indent=1)
    print({k: (len(v["steps"]), len(v["checks"])) for k, v in out["scripts"].items()}, len(out["showdowns"]))


if __name__ == "__main__":
    main()
