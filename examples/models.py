# -*- coding: utf-8 -*-
"""
The reference's example models, re-expressed against ``popdynamics_b200.basepop``.

These are the workloads BASELINE.json names (SIR, SEIR +- demography, two-strain SIR, TB, RMIT TB,
and the SIR of mix_model.py).  Parameter names, default values, flow registration order and the
arithmetic order inside ``calculate_vars`` follow the reference example scripts (cited per class,
paths relative to /root/reference) so that trajectories are comparable number for number; the
plotting / main-routine parts of those scripts are out of scope.  Constructors accept 1-D numpy
arrays for swept parameters (one value per ensemble member).
"""

import math

import numpy

from popdynamics_b200 import basepop
from popdynamics_b200.basepop import old_div


def _ceil(x):
    return numpy.ceil(x) if isinstance(x, numpy.ndarray) else math.ceil(x)


def _floor(x):
    return numpy.floor(x) if isinstance(x, numpy.ndarray) else math.floor(x)


class _IncidencePrevalenceMixin(object):
    """Diagnostics shared by the SIR-type examples (examples/sir_model.py:144-160)."""

    def calculate_diagnostic_vars(self):
        self.vars["incidence"] = 0.
        for from_label, to_label, rate in self.var_transfer_rate_flows:
            val = self.compartments[from_label] * self.vars[rate]
            if "infectious" in to_label:
                self.vars["incidence"] += old_div(val, self.vars["population"])
        self.vars["prevalence"] = old_div(self.compartments["infectious"],
                                          self.vars["population"])


class SirModel(_IncidencePrevalenceMixin, basepop.BaseModel):
    """SIR, C=3 (examples/sir_model.py:51-160; BASELINE config 1)."""

    defaults = (("population", 50), ("start_infectious", 1.), ("r0", 12.),
                ("duration_preinfectious", 8.), ("duration_infectious", 7.),
                ("life_expectancy", 70. * 365.))
    population_key = "population"

    def __init__(self, params={}):
        basepop.BaseModel.__init__(self)
        for key, value in self.defaults:
            self.set_param(key, value)
        for key, value in params.items():
            self.set_param(key, value)
        p = self.params
        self.set_compartment("susceptible", p[self.population_key] - p["start_infectious"])
        self.set_compartment("infectious", p["start_infectious"])
        self.set_compartment("immune", 0.)
        self.set_param("infection_beta",
                       p["r0"] / (p["duration_infectious"] * p[self.population_key]))
        self.set_param("infection_rate_recover", 1. / p["duration_infectious"])

    def set_flows(self):
        self.set_var_transfer_rate_flow("susceptible", "infectious", "rate_force")
        self.set_fixed_transfer_rate_flow("infectious", "immune", "infection_rate_recover")

    def calculate_vars(self):
        self.vars["population"] = sum(self.compartments.values())
        self.vars["rate_force"] = self.params["infection_beta"] * self.compartments["infectious"]


class MixSirModel(SirModel):
    """The SIR variant of examples/mix_model.py:25-103 (population 1000, r0 5)."""

    defaults = (("start_population", 1000), ("start_infectious", 1.), ("r0", 5.),
                ("duration_preinfectious", 8.), ("duration_infectious", 7.),
                ("life_expectancy", 70. * 365.))
    population_key = "start_population"


class SeirModel(basepop.BaseModel):
    """SEIR, C=4 (examples/seir_model.py:49-145)."""

    def __init__(self, params):
        basepop.BaseModel.__init__(self)
        for key in params:
            self.params[key] = params[key]
        p = self.params
        self.set_compartment('susceptible', p['population'] - p['start_infectious'])
        self.set_compartment('preinfectious', 0.0)
        self.set_compartment('infectious', p['start_infectious'])
        self.set_compartment('immune', 0.0)
        self.set_param('infection_beta',
                       1.0 * p['r0'] / (p['duration_infectious'] * p['population']))
        self.set_param('infection_rate_progress', 1.0 / p['duration_preinfectious'])
        self.set_param('infection_rate_recover', 1.0 / p['duration_infectious'])

    def set_flows(self):
        self.set_var_transfer_rate_flow('susceptible', 'preinfectious', 'rate_force')
        self.set_fixed_transfer_rate_flow('preinfectious', 'infectious',
                                          'infection_rate_progress')
        self.set_fixed_transfer_rate_flow('infectious', 'immune', 'infection_rate_recover')

    def calculate_vars(self):
        self.vars['population'] = sum(self.compartments.values())
        self.vars['rate_force'] = self.params['infection_beta'] * self.compartments['infectious']

    def calculate_diagnostic_vars(self):
        population = self.vars['population']
        self.vars['incidence'] = 0.0
        for from_label, to_label, rate in self.fixed_transfer_rate_flows:
            val = self.compartments[from_label] * rate
            if 'infectious' in to_label:
                self.vars['incidence'] += val / population
        self.vars['infections'] = 0.
        for from_label, to_label, rate in self.var_transfer_rate_flows:
            val = self.compartments[from_label] * self.vars[rate]
            if 'preinfectious' in to_label:
                self.vars['infections'] += val / population
        self.vars['prevalence'] = self.compartments['infectious'] / population


class SeirDemographyModel(SeirModel):
    """SEIR with births and background deaths (examples/seir_model.py:147-228; config 2)."""

    def __init__(self, param_dictionary):
        basepop.BaseModel.__init__(self)
        d = param_dictionary
        self.set_compartment('susceptible', d['population'] - d['start_infectious'])
        self.set_compartment('preinfectious', 0.)
        self.set_compartment('infectious', d['start_infectious'])
        self.set_compartment('immune', 0.)
        self.set_param('r0', d['r0'])
        self.set_param('infection_beta', d['r0'] / (d['duration_infectious'] * d['population']))
        self.set_param('infection_rate_progress', old_div(1., d['duration_preinfectious']))
        self.set_param('infection_rate_recover', old_div(1., d['duration_infectious']))
        self.set_param('demo_rate_death', old_div(1., d['life_expectancy']))
        self.set_param('demo_rate_birth', self.params['demo_rate_death'])

    def set_flows(self):
        self.set_var_entry_rate_flow('susceptible', 'rate_birth')
        self.set_var_transfer_rate_flow('susceptible', 'preinfectious', 'rate_force')
        self.set_fixed_transfer_rate_flow('preinfectious', 'infectious',
                                          'infection_rate_progress')
        self.set_fixed_transfer_rate_flow('infectious', 'immune', 'infection_rate_recover')
        self.set_background_death_rate('demo_rate_death')

    def calculate_vars(self):
        self.vars['population'] = sum(self.compartments.values())
        self.vars['rate_birth'] = self.params['demo_rate_birth'] * self.vars['population']
        self.vars['rate_force'] = self.params['infection_beta'] * self.compartments['infectious']


SEIR_PARAMS = {
    'measles': {'population': 1e6, 'start_infectious': 1., 'r0': 13.,
                'duration_preinfectious': 8., 'duration_infectious': 7.,
                'life_expectancy': 70. * 365.},
    'flu': {'population': 1e6, 'start_infectious': 1., 'r0': 2.,
            'duration_preinfectious': 2., 'duration_infectious': 2.,
            'life_expectancy': 70. * 365.},
}   # examples/seir_model.py:391-408


class StrainsModel(basepop.BaseModel):
    """Two competing SIR strains, C=5 (examples/stochastic_strains_model.py:31-151; the
    headline config 3).  Starts at the resident-strain equilibrium with one invader."""

    def __init__(self, input_params=[]):
        basepop.BaseModel.__init__(self)
        for key, value in (("r0_resident", 2), ("r0_invader", 3), ("rate_birth", 2),
                           ("rate_death", 1.0e-3), ("rate_recover", 1.0e-2),
                           ("rate_infection_death", 0)):
            self.params[key] = value
        for key, value in input_params:
            self.params[key] = value
        p = self.params
        self.set_param("s0", old_div(p["rate_birth"], p["rate_death"]))
        for strain in ("resident", "invader"):
            self.set_param(
                "beta_" + strain,
                p["r0_" + strain] * (p["rate_recover"] + p["rate_infection_death"]
                                     + p["rate_death"]) / p["s0"])
        self.set_compartment("susceptible", _ceil(old_div(p["s0"], p["r0_resident"])))
        self.set_compartment(
            "infectious_resident",
            _floor(p["rate_death"] / p["beta_resident"] * (p["r0_resident"] - 1)))
        self.set_compartment("infectious_invader", 1)
        self.set_compartment("recovered_resident", 0)
        self.set_compartment("recovered_invader", 0)

    def calculate_vars(self):
        self.vars["population"] = sum(self.compartments.values())
        self.vars["rate_birth"] = self.params["rate_birth"]
        self.vars["rate_infection_invader"] = \
            self.compartments["infectious_invader"] * self.params["beta_invader"]
        self.vars["rate_infection_resident"] = \
            self.compartments["infectious_resident"] * self.params["beta_resident"]

    def set_flows(self):
        self.set_var_entry_rate_flow("susceptible", "rate_birth")
        self.set_background_death_rate("rate_death")
        self.set_infection_death_rate_flow("infectious_invader", "rate_infection_death")
        self.set_infection_death_rate_flow("infectious_resident", "rate_infection_death")
        self.set_var_transfer_rate_flow("susceptible", "infectious_invader",
                                        "rate_infection_invader")
        self.set_var_transfer_rate_flow("susceptible", "infectious_resident",
                                        "rate_infection_resident")
        self.set_fixed_transfer_rate_flow("infectious_resident", "recovered_resident",
                                          "rate_recover")
        self.set_fixed_transfer_rate_flow("infectious_invader", "recovered_invader",
                                          "rate_recover")

    def calculate_diagnostic_vars(self):
        population = self.vars["population"]
        self.vars["incidence"] = 0.
        for from_label, to_label, rate in self.var_transfer_rate_flows:
            val = self.compartments[from_label] * self.vars[rate]
            if "infectious" in to_label:
                self.vars["incidence"] += old_div(val, population)
        self.vars["prevalence"] = 0.
        for label, val in list(self.compartments.items()):
            if "infectious" in label:
                self.vars["prevalence"] += old_div(val, population)


class SimpleTbModel(basepop.BaseModel):
    """AuTuMN-style TB model, C=6 (7 with BCG) with two sigmoidal scale-ups
    (examples/tb_model.py:41-177; config 4)."""

    def __init__(self, interventions=[]):
        basepop.BaseModel.__init__(self)
        time_treatment = .5
        rates = {
            'demo_rate_birth': old_div(20., 1e3),
            'demo_rate_death': old_div(1., 65),
            'tb_n_contact': 40.,
            'tb_rate_earlyprogress': old_div(.1, .5),
            'tb_rate_lateprogress': old_div(.1, 100.),
            'tb_rate_stabilise': old_div(.9, .5),
            'tb_rate_recover': old_div(.6, 3.),
            'tb_rate_death': old_div(.4, 3.),
            'program_rate_completion_infect': old_div(.9, time_treatment),
            'program_rate_default_infect': old_div(.05, time_treatment),
            'program_rate_death_infect': old_div(.05, time_treatment),
            'program_rate_completion_noninfect': old_div(.9, time_treatment),
            'program_rate_default_noninfect': old_div(.05, time_treatment),
            'program_rate_death_noninfect': old_div(.05, time_treatment),
        }
        for key in rates:
            self.params[key] = rates[key]
        self.interventions = interventions
        for label in ('susceptible', 'latent_early', 'latent_late', 'active',
                      'treatment_infect', 'treatment_noninfect'):
            self.set_compartment(label, 0.)
        self.set_compartment('susceptible', 1e6)
        self.set_compartment('active', 1.)
        if 'vaccination' in self.interventions:
            self.set_compartment('susceptible_vaccinated', 0.)
        early = basepop.make_sigmoidal_curve(y_high=2, y_low=0, x_start=1950, x_inflect=1970,
                                             multiplier=4)
        late = basepop.make_sigmoidal_curve(y_high=4, y_low=2, x_start=1995, x_inflect=2003,
                                            multiplier=3)
        self.set_scaleup_fn('program_rate_detect', lambda x: early(x) if x < 1990 else late(x))
        if 'vaccination' in self.interventions:
            self.set_param('int_vaccine_efficacy', .5)
            self.set_scaleup_fn('int_vaccination', basepop.make_sigmoidal_curve(
                y_high=.95, y_low=0, x_start=1921, x_inflect=2006, multiplier=3))

    def calculate_vars(self):
        v, p = self.vars, self.params
        v['population'] = sum(self.compartments.values())
        v['rate_birth'] = p['demo_rate_birth'] * v['population']
        v['infectious_population'] = 0.
        for label in self.labels:
            if 'active' in label or '_infect' in label:
                v['infectious_population'] += self.compartments[label]
        v['rate_force'] = p['tb_n_contact'] * v['infectious_population'] / v['population']
        if 'vaccination' in self.interventions:
            v['rate_birth_vaccinated'] = v['rate_birth'] * v['int_vaccination']
            v['rate_birth'] -= v['rate_birth_vaccinated']
            v['rate_force_vaccinated'] = v['rate_force'] * p['int_vaccine_efficacy']

    def set_flows(self):
        self.set_var_entry_rate_flow('susceptible', 'rate_birth')
        self.set_background_death_rate('demo_rate_death')
        self.set_var_transfer_rate_flow('susceptible', 'latent_early', 'rate_force')
        self.set_fixed_transfer_rate_flow('latent_early', 'active', 'tb_rate_earlyprogress')
        self.set_fixed_transfer_rate_flow('latent_early', 'latent_late', 'tb_rate_stabilise')
        self.set_fixed_transfer_rate_flow('latent_late', 'active', 'tb_rate_lateprogress')
        self.set_fixed_transfer_rate_flow('active', 'latent_late', 'tb_rate_recover')
        self.set_infection_death_rate_flow('active', 'tb_rate_death')
        self.set_var_transfer_rate_flow('active', 'treatment_infect', 'program_rate_detect')
        self.set_fixed_transfer_rate_flow('treatment_infect', 'treatment_noninfect',
                                          'program_rate_completion_infect')
        self.set_fixed_transfer_rate_flow('treatment_infect', 'active',
                                          'program_rate_default_infect')
        self.set_fixed_transfer_rate_flow('treatment_noninfect', 'susceptible',
                                          'program_rate_completion_noninfect')
        self.set_fixed_transfer_rate_flow('treatment_noninfect', 'active',
                                          'program_rate_default_noninfect')
        self.set_infection_death_rate_flow('treatment_infect', 'program_rate_death_infect')
        self.set_infection_death_rate_flow('treatment_noninfect',
                                           'program_rate_death_noninfect')
        if 'vaccination' in self.interventions:
            self.set_var_entry_rate_flow('susceptible_vaccinated', 'rate_birth_vaccinated')
            self.set_var_transfer_rate_flow('susceptible_vaccinated', 'latent_early',
                                            'rate_force_vaccinated')

    def calculate_diagnostic_vars(self):
        v = self.vars
        v['prevalence'] = v['infectious_population'] / v['population'] * 1e5
        v['mortality'] = v['rate_infection_death'] / v['population'] * 1e5
        rate_incidence = 0.
        for from_label, to_label, rate in self.fixed_transfer_rate_flows:
            val = self.compartments[from_label] * rate
            if 'latent' in from_label and 'active' in to_label:
                rate_incidence += val
        v['incidence'] = rate_incidence / v['population'] * 1e5
        v['latent'] = 0.
        for label in self.labels:
            if 'latent' in label:
                v['latent'] += (self.compartments[label] / v['population'] * 1e5)


class ScaleupTbModel(SimpleTbModel):
    """Declared extension for BASELINE config 4: the reference's stochastic loops
    never evaluate scale-up functions, so ``SimpleTbModel`` raises KeyError there; this subclass
    evaluates them at the start of ``calculate_vars`` -- only usable where ``self.time`` is a
    plain number (host diagnostics / the reference itself)."""

    def calculate_vars(self):
        self.calculate_scaleup_vars()
        SimpleTbModel.calculate_vars(self)


class RmitTbModel(basepop.BaseModel):
    """Mwangi's TB model on population fractions, C=6 (examples/rmit_tb_model.py:23-102;
    config 5 sweeps beta / rho / sigma1..3)."""

    def __init__(self, parameters):
        basepop.BaseModel.__init__(self)
        for parameter, value in parameters.items():
            self.set_param(parameter, value)
        for label in ('S', 'L1', 'L2', 'P', 'I', 'R'):
            self.set_compartment(label, 0.)
        self.set_compartment('S', 1.)
        self.set_compartment('I', 1e-3)
        p = self.params
        self.set_param('f_phi', p['f'] * p['phi'])
        self.set_param('1-f_phi', (1. - p['f']) * p['phi'])
        self.set_param('tau_alpha', p['tau'] + p['alpha'])

    def calculate_vars(self):
        v, p, c = self.vars, self.params, self.compartments
        v['population'] = sum(c.values())
        v['lambda'] = p['mu'] * v['population'] + p['d'] * c['I']
        v['beta_I'] = p['beta'] * c['I'] / v['population']
        for i_sigma in range(1, 4):
            v['sigma%d_beta_I' % i_sigma] = v['beta_I'] * p['sigma%d' % i_sigma]

    def set_flows(self):
        self.set_var_entry_rate_flow('S', 'lambda')
        self.set_background_death_rate('mu')
        self.set_var_transfer_rate_flow('S', 'L1', 'beta_I')
        self.set_var_transfer_rate_flow('L2', 'L1', 'sigma1_beta_I')
        self.set_var_transfer_rate_flow('P', 'L1', 'sigma2_beta_I')
        self.set_var_transfer_rate_flow('R', 'L1', 'sigma3_beta_I')
        self.set_fixed_transfer_rate_flow('L1', 'I', 'f_phi')
        self.set_fixed_transfer_rate_flow('L1', 'L2', '1-f_phi')
        self.set_fixed_transfer_rate_flow('L2', 'I', 'eta')
        self.set_fixed_transfer_rate_flow('L1', 'P', 'theta')
        self.set_fixed_transfer_rate_flow('L2', 'P', 'rho')
        self.set_fixed_transfer_rate_flow('I', 'R', 'tau_alpha')
        self.set_infection_death_rate_flow('I', 'd')
        self.set_fixed_transfer_rate_flow('R', 'I', 'omega')

    def calculate_diagnostic_vars(self):
        self.vars['proportion'] = old_div(self.compartments['I'], self.vars['population'])


RMIT_FIXED_PARAMS = {
    'mu': old_div(1., 70.), 'd': .1, 'phi': 12, 'f': .05, 'eta': 2e-4, 'tau': old_div(1., 2.),
    'alpha': old_div(2., 9.), 'theta': 1., 'rho': 0.1, 'omega': 2e-5,
    'sigma1': 0., 'sigma2': 0., 'sigma3': 0.,
}   # examples/rmit_tb_model.py:114-131


MODEL_CLASSES = {
    'SirModel': SirModel, 'MixSirModel': MixSirModel, 'SeirModel': SeirModel,
    'SeirDemographyModel': SeirDemographyModel, 'StrainsModel': StrainsModel,
    'SimpleTbModel': SimpleTbModel, 'RmitTbModel': RmitTbModel,
}
