import math


class _HP(object):
    def __init__(self, name, lower=None, upper=None, log=False, default_value=None, choices=None):
        self.name = name
        self.lower = lower
        self.upper = upper
        self.log = log
        self.default_value = default_value
        self.choices = choices


class UniformFloatHyperparameter(_HP):
    def _from_unit(self, u):
        if self.log:
            lo, hi = math.log(self.lower), math.log(self.upper)
            return float(math.exp(lo + (hi - lo) * u))
        return float(self.lower + (self.upper - self.lower) * u)


class UniformIntegerHyperparameter(_HP):
    def _from_unit(self, u):
        if self.log:
            lo, hi = math.log(self.lower - 0.49999), math.log(self.upper + 0.49999)
            v = math.exp(lo + (hi - lo) * u)
        else:
            v = (self.lower - 0.49999) + ((self.upper + 0.49999) - (self.lower - 0.49999)) * u
        return int(min(max(int(round(v)), self.lower), self.upper))


class CategoricalHyperparameter(_HP):
    def __init__(self, name, choices, default_value=None):
        super(CategoricalHyperparameter, self).__init__(name, choices=list(choices), default_value=default_value)

    def _from_unit(self, u):
        return self.choices[min(int(u * len(self.choices)), len(self.choices) - 1)]
