"""Committed golden vectors (tests/golden/oracle_pins.json, recipe in tests/pins.py): the oracle must reproduce them on
the CPU, and the device library must reproduce them through the C ABI with no oracle in the loop."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLDEN = json.load(open(os.path.join(HERE, "golden", "oracle_pins.json")))


def _compare(got):
    for key, want in got.items():
        assert GOLDEN[key] == want, key


def test_oracle_reproduces_the_committed_pins(oracle):
    import pins

    _compare(pins.compute(pins.OracleBackend(oracle)))
    _compare(pins.contract_pins(oracle))


@pytest.mark.gpu
def test_device_reproduces_the_committed_pins(rbp):
    import pins

    _compare(pins.compute(pins.DeviceBackend(rbp)))
