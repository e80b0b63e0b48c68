class Concentration(object):
    def __init__(self, *a, mass_def=None, **k):
        self.mass_def = mass_def
