"""Sweep-axis iterators, same cumulative recurrences as the reference
(ahkab/utilities.py:246-378) so that sweep / frequency grids are bit-identical
(values are running products / sums, not linspace/logspace)."""
import numpy as np


class log_axis_iterator(object):
    """ahkab/utilities.py:246-304."""
    def __init__(self, min, max, points):
        self.inc = 10. ** ((np.log10(max) - np.log10(min)) / (points - 1))
        self.max, self.min = max, min
        self.index = 0
        self.current = min
        self.points = points

    def __next__(self):
        if self.index == 0:
            ret = self.current
        elif self.index < self.points:
            self.current = self.current * self.inc
            ret = self.current
        else:
            raise StopIteration
        self.index = self.index + 1
        return ret

    next = __next__

    def __getitem__(self, i):
        if i == 0:
            return self.min
        elif i < self.points:
            return self.min * self.inc ** i
        return None

    def __iter__(self):
        return self


class lin_axis_iterator(object):
    """ahkab/utilities.py:307-378."""
    def __init__(self, min, max, points):
        if points < 2 and min != max:
            raise ValueError('Linear iterator from %d to %d with %d points.' %
                             (min, max, points))
        if points < 1:
            raise ValueError('Linear iterator from %d to %d with %d points.' %
                             (min, max, points))
        if points > 1:
            self.inc = (max - min) / (points - 1)
        elif points == 1 and max == min:
            self.inc = 0
        self.max, self.min = max, min
        self.index = 0
        self.current = min
        self.points = points

    def __next__(self):
        if self.index == 0:
            pass
        elif self.index < self.points:
            self.current = self.current + self.inc
        else:
            raise StopIteration
        ret = self.current
        self.index = self.index + 1
        return ret

    next = __next__

    def __getitem__(self, i):
        if i < self.points:
            return self.min + self.inc * i
        return None

    def __iter__(self):
        return self
