"""ConfigSpace-free spaces (tlbo_b200/config_space/space.py): array encoding, get_types, default imputation,
conditions -> has_conditions, and the JSON-lines history loader, on the reference's own random-forest and lda
search spaces (tlbo/config_space/space_instance.py:11-35, :98)."""
import json

import numpy as np

from tlbo_b200.config_space import get_types
from tlbo_b200.config_space import space as S


def rf_space():
    cs = S.ConfigurationSpace()
    cs.add_hyperparameters([
        S.CategoricalHyperparameter('criterion', ['gini', 'entropy'], default_value='gini'),
        S.UniformFloatHyperparameter('max_features', 0, 1, default_value=0.5, q=0.05),
        S.UnParametrizedHyperparameter('max_depth', 'None'),
        S.UniformIntegerHyperparameter('min_samples_split', 2, 20, default_value=2),
        S.UniformIntegerHyperparameter('min_samples_leaf', 1, 20, default_value=1),
        S.Constant('min_weight_fraction_leaf', 0.),
        S.UnParametrizedHyperparameter('max_leaf_nodes', 'None'),
        S.CategoricalHyperparameter('bootstrap', ['True', 'False'], default_value='True'),
        S.UnParametrizedHyperparameter('min_impurity_decrease', 0.)])
    return cs


def test_random_forest_space_types_and_encoding():
    cs = rf_space()
    names = cs.get_hyperparameter_names()
    assert names == sorted(names)                       # no conditions: alphabetical
    types, bounds = get_types(cs)
    # bootstrap, criterion categorical (2 choices); everything else numerical / constant: the RF_TYPES of synthetic.py
    assert types.tolist() == [2, 2, 0, 0, 0, 0, 0, 0, 0] and not types.has_conditions
    c = S.Configuration(cs, values=dict(criterion='entropy', max_features=0.25, max_depth='None', min_samples_split=20,
                                        min_samples_leaf=1, min_weight_fraction_leaf=0., max_leaf_nodes='None',
                                        bootstrap='False', min_impurity_decrease=0.))
    a = c.get_array()
    assert a[names.index('bootstrap')] == 1.0 and a[names.index('criterion')] == 1.0
    assert a[names.index('max_features')] == 0.25 and a[names.index('max_depth')] == 0.0
    assert abs(a[names.index('min_samples_split')] - (20 - 1.50001) / (20.49999 - 1.50001)) < 1e-15
    assert abs(a[names.index('min_samples_leaf')] - (1 - 0.50001) / (20.49999 - 0.50001)) < 1e-15
    back = S.Configuration(cs, vector=a).get_dictionary()
    assert back['min_samples_split'] == 20 and back['criterion'] == 'entropy' and back['max_features'] == 0.25
    assert S.Configuration(cs, vector=a) == c and hash(S.Configuration(cs, vector=a)) == hash(c)
    d = cs.get_default_configuration().get_array()
    assert d[names.index('max_features')] == 0.5 and d[names.index('bootstrap')] == 0.0


def test_conditional_space_activity_imputation_and_flag(tmp_path):
    # lda: shrinkage_factor is active only for shrinkage == 'manual' (space_instance.py:98-110)
    cs = S.ConfigurationSpace()
    shrinkage = S.CategoricalHyperparameter('shrinkage', ['None', 'auto', 'manual'], default_value='None')
    factor = S.UniformFloatHyperparameter('shrinkage_factor', 0., 1., 0.5)
    n_comp = S.UniformIntegerHyperparameter('n_components', 1, 250, default_value=10)
    tol = S.UniformFloatHyperparameter('tol', 1e-5, 1e-1, default_value=1e-4, log=True)
    cs.add_hyperparameters([shrinkage, factor, n_comp, tol])
    cs.add_condition(S.EqualsCondition(factor, shrinkage, 'manual'))
    names = cs.get_hyperparameter_names()
    assert names == ['n_components', 'shrinkage', 'tol', 'shrinkage_factor']     # the child after its parent's level
    types, _ = get_types(cs)
    assert types.tolist() == [0, 3, 0, 0] and types.has_conditions
    a = S.Configuration(cs, values=dict(shrinkage='auto', shrinkage_factor=0.3, n_components=10, tol=1e-3))
    b = S.Configuration(cs, values=dict(shrinkage='manual', shrinkage_factor=0.3, n_components=10, tol=1e-3))
    assert np.isnan(a.get_array()[3]) and b.get_array()[3] == 0.3
    assert abs(a.get_array()[2] - 0.5) < 1e-12                                   # log scale: 1e-3 is half way
    X = S.convert_configurations_to_array([a, b])
    assert X[0, 3] == 0.5 and X[1, 3] == 0.3                                     # inactive -> normalised default
    # history round trip through the JSON-lines format of tools/export_history.py
    p = tmp_path / 'h.jsonl'
    with open(p, 'w') as f:
        for c, perf in ((a, 0.25), (b, 0.125)):
            f.write(json.dumps({'config': c.get_dictionary(), 'perf': perf}) + '\n')
    hist = S.load_history(str(p), cs)
    assert list(hist.values()) == [0.25, 0.125] and list(hist.keys()) == [a, b]
