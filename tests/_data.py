"""Test helper: rebuild (drug, channel) datasets from tests/golden/datasets.npz with the grouping
semantics of the reference loader (python/doseresponse.py:31-37, 60-67)."""
import itertools as it
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _unique_in_order(a):
    _, idx = np.unique(a, return_index=True)
    return a[np.sort(idx)]


class Table:
    def __init__(self, name="crumb_data"):
        z = np.load(os.path.join(GOLD, "datasets.npz"))
        self.drug = z[name + "__drug"]
        self.channel = z[name + "__channel"]
        self.experiment = z[name + "__experiment"]
        self.dose = z[name + "__dose"]
        self.response = z[name + "__response"]
        self.drugs = _unique_in_order(self.drug)
        self.channels = _unique_in_order(self.channel)

    def pairs(self):
        return list(it.product(self.drugs, self.channels))

    def experiments(self, drug, channel):
        m = (self.drug == drug) & (self.channel == channel)
        out = []
        for e in _unique_in_order(self.experiment[m]):
            mm = m & (self.experiment == e)
            out.append(np.stack([self.dose[mm], self.response[mm]], axis=1))
        return out

    def concat(self, drug, channel):
        ex = self.experiments(drug, channel)
        concs = np.concatenate([e[:, 0] for e in ex])
        responses = np.concatenate([e[:, 1] for e in ex])
        return concs, responses
