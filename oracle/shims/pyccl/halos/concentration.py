from .halo_model_base import Concentration


def _mk(name):
    return type(name, (Concentration,), {})


ConcentrationDuffy08 = _mk('ConcentrationDuffy08')
ConcentrationKlypin11 = _mk('ConcentrationKlypin11')
ConcentrationPrada12 = _mk('ConcentrationPrada12')
ConcentrationDiemer15 = _mk('ConcentrationDiemer15')
ConcentrationIshiyama21 = _mk('ConcentrationIshiyama21')
ConcentrationBhattacharya13 = _mk('ConcentrationBhattacharya13')
ConcentrationConstant = _mk('ConcentrationConstant')
