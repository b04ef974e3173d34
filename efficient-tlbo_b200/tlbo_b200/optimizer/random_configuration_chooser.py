"""Random-interleaving rule used by the offline loop
(reference tlbo/optimizer/random_configuration_chooser.py ``ChooserProb``)."""


class ChooserProb(object):
    def __init__(self, prob, rng):
        self.prob = prob
        self.rng = rng

    def next_smbo_iteration(self):
        pass

    def check(self, iteration):
        return self.rng.rand() < self.prob
