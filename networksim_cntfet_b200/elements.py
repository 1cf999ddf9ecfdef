"""Conduction elements (mirror of cnet.py:20-87).

The classes keep the reference's names, constructor arguments and `get_conductance`
behaviour so that `element=LinExpTransistor` keeps working as a constructor argument
(netsim.py:27, measure_perc.py:23-25).  On the GPU path they are never instantiated per
junction: `conductance_table` evaluates the same formula once per (junction kind, gate level)
on the host -- with numpy, exactly as the reference does -- and the assembly kernel looks the
value up, so device conductances are bit-identical to the reference's.
"""
import numpy as np

KIND_CHARS = "vms"           # u8 codes: 0 electrode, 1 metallic, 2 semiconducting


def pair_code(kind):
    """'ms' -> 3*1+2"""
    return 3 * KIND_CHARS.index(kind[0]) + KIND_CHARS.index(kind[1])


def pair_string(code):
    return KIND_CHARS[code // 3] + KIND_CHARS[code % 3]


class LinExpTransistor(object):
    """cnet.py:20-41.  G = exp(-alpha*Vg) * exp(-10*alpha), alpha chosen by junction kind."""
    onoffmappings = [
        {'ms': 0.5, 'sm': 0.5, 'mm': 1e-5, 'ss': 1e-5, 'vs': 1e-5, 'sv': 1e-5, 'vm': 1e-5, 'mv': 1e-5},
        {'ms': 0.5, 'sm': 0.5, 'mm': 1e-5, 'ss': 1e-5, 'vs': 0.5, 'sv': 0.5, 'vm': 1e-5, 'mv': 1e-5},
        {'ms': 0.5, 'sm': 0.5, 'mm': 1e-5, 'ss': 0.5, 'vs': 0.5, 'sv': 0.5, 'vm': 1e-5, 'mv': 1e-5},
    ]

    def __init__(self, type, onoffmap=0):
        self.gate_voltage = 0
        self.type = type
        self.alpha = self.onoffmappings[onoffmap][type]

    def lin_exp(self, vg):
        alpha = self.alpha
        normalization = np.exp(-10 * alpha)
        return np.exp(-alpha * vg) * normalization

    def get_conductance(self):
        return self.lin_exp(self.gate_voltage)


class FermiDiracTransistor(object):
    """cnet.py:45-66.  Fermi-Dirac step in Vg; only onoffmap 0 exists."""
    onoffmappings = [
        {'ms': 1 / 5e-5, 'sm': 1 / 5e-5, 'mm': 1, 'ss': 1, 'vs': 1, 'sv': 1, 'vm': 1, 'mv': 1},
    ]

    def __init__(self, type, onoffmap=0):
        self.offG = 1 / self.onoffmappings[onoffmap][type]
        self.scaling = 1 - self.offG
        self.offset = self.offG
        self.gate_voltage = 0

    def _fermi_dirac(self, x, scaling, offset, threshold):
        return scaling * (1 / (np.exp(10 * (x - threshold)) + 1)) + offset

    def get_conductance(self, gate=False):
        if gate:
            return self._fermi_dirac(gate, self.scaling, self.offset, 0)
        return self._fermi_dirac(self.gate_voltage, self.scaling, self.offset, 0)


class StepTransistor(object):
    """cnet.py:67-79 (unused by any caller of the hot path; kept for API completeness)."""

    def __init__(self, on_resistance=1, off_resistance=1000, threshold_voltage=0, gate_voltage=0):
        assert on_resistance > 0 and off_resistance > 0, "ERROR: a component cannot have -ve resistance"
        self.on_resistance = on_resistance
        self.off_resistance = off_resistance
        self.threshold_voltage = threshold_voltage
        self.gate_voltage = gate_voltage

    def get_conductance(self):
        if self.gate_voltage <= self.threshold_voltage:
            return 1 / self.on_resistance
        return 1 / self.off_resistance


class Resistor(object):
    """cnet.py:81-87."""

    def __init__(self, R=1):
        assert R > 0, "ERROR: a component cannot have -ve resistance"
        self.resistance = R
        self.conductance = 1 / R

    def get_conductance(self):
        return self.conductance


def conductance_table(element, onoffmap, gate_levels):
    """[9, nlevels] float64: conductance of every junction kind at every gate level, computed
    by instantiating the reference-compatible element and calling get_conductance(), i.e. the
    very arithmetic of cnet.py:99-103.  Kinds the element has no entry for ('vv') are NaN."""
    levels = [float(v) for v in gate_levels]
    tab = np.full((9, len(levels)), np.nan, dtype=np.float64)
    for code in range(9):
        try:
            comp = element(pair_string(code), onoffmap)
        except KeyError:
            continue
        for k, vg in enumerate(levels):
            comp.gate_voltage = vg
            with np.errstate(over="ignore"):
                tab[code, k] = comp.get_conductance()
    return tab
