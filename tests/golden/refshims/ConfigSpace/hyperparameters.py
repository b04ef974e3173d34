class _HP(object):
    def __init__(self, name, normalized_default_value=0.0):
        self.name = name
        self.normalized_default_value = normalized_default_value


class CategoricalHyperparameter(_HP):
    def __init__(self, name, choices, default_value=None):
        super().__init__(name, 0.0)
        self.choices = list(choices)


class OrdinalHyperparameter(_HP):
    def __init__(self, name, sequence, default_value=None):
        super().__init__(name, 0.0)
        self.sequence = list(sequence)


class Constant(_HP):
    def __init__(self, name, value):
        super().__init__(name, 0.0)
        self.value = value


class UniformFloatHyperparameter(_HP):
    def __init__(self, name, lower, upper, default_value=None, q=None, log=False):
        super().__init__(name, 0.5)
        self.lower, self.upper, self.log, self.q = lower, upper, log, q


class UniformIntegerHyperparameter(_HP):
    def __init__(self, name, lower, upper, default_value=None, q=None, log=False):
        super().__init__(name, 0.5)
        self.lower, self.upper, self.log, self.q = lower, upper, log, q
