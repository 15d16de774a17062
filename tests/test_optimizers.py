"""The Keras (OptimizerV2, TF 2.x defaults) update rules of bear_b200._engine.Optimizer other than Adam -- bear_net.train
takes any ``tf.keras.optimizers`` name (bear_net.py:264) -- against independent numpy restatements of the published
formulas.  The reference pins none of them; Adam is covered by tests/test_oracle_pinned.py and the GPU trajectory test."""
import numpy as np
import pytest
import torch


def _numpy_rule(name, lr):
    st = {'t': 0, 'ms': 1.0}

    def step(p, g):
        st['t'] += 1
        t = st['t']
        if name == 'SGD':
            return p - lr * g
        if name == 'RMSprop':                                   # rho 0.9, epsilon 1e-7
            st['v'] = 0.9 * st.get('v', 0.0) + 0.1 * g * g
            return p - lr * g / (np.sqrt(st['v']) + 1e-7)
        if name == 'Adagrad':                                   # initial accumulator 0.1
            st['v'] = st.get('v', 0.1) + g * g
            return p - lr * g / (np.sqrt(st['v']) + 1e-7)
        if name == 'Adadelta':                                  # rho 0.95
            st['v'] = 0.95 * st.get('v', 0.0) + 0.05 * g * g
            upd = g * np.sqrt(st.get('d', 0.0) + 1e-7) / np.sqrt(st['v'] + 1e-7)
            st['d'] = 0.95 * st.get('d', 0.0) + 0.05 * upd * upd
            return p - lr * upd
        if name == 'Adamax':
            st['m'] = 0.9 * st.get('m', 0.0) + 0.1 * g
            st['u'] = np.maximum(0.999 * st.get('u', 0.0), np.abs(g))
            return p - lr / (1 - 0.9 ** t) * st['m'] / (st['u'] + 1e-7)
        if name == 'Nadam':
            u_t = 0.9 * (1 - 0.5 * 0.96 ** (0.004 * t))
            u_n = 0.9 * (1 - 0.5 * 0.96 ** (0.004 * (t + 1)))
            st['ms'] *= u_t
            st['m'] = 0.9 * st.get('m', 0.0) + 0.1 * g
            st['v'] = 0.999 * st.get('v', 0.0) + 0.001 * g * g
            g_hat, m_hat = g / (1 - st['ms']), st['m'] / (1 - st['ms'] * u_n)
            v_hat = st['v'] / (1 - 0.999 ** t)
            return p - lr * ((1 - u_t) * g_hat + u_n * m_hat) / (np.sqrt(v_hat) + 1e-7)
        raise KeyError(name)
    return step


@pytest.mark.parametrize('name', ['SGD', 'RMSprop', 'Adagrad', 'Adadelta', 'Adamax', 'Nadam'])
def test_keras_update_rules(name):
    from bear_b200 import _engine as eng
    rng = np.random.default_rng(3)
    p0 = rng.normal(size=17)
    opt = eng.Optimizer(name, 0.01, 17, 'cpu')
    rule = _numpy_rule(name, 0.01)
    p_t, p_n = torch.from_numpy(p0.copy()), p0.copy()
    for _ in range(25):
        g = rng.normal(size=17) * rng.choice([1e-3, 1.0, 30.0])
        opt.apply(p_t, torch.from_numpy(g))
        p_n = rule(p_n, g)
    assert np.allclose(p_t.numpy(), p_n, rtol=1e-12, atol=1e-15)


def test_unknown_optimizer_is_an_error():
    from bear_b200 import _engine as eng
    with pytest.raises(ValueError):
        eng.Optimizer('Ftrl', 0.01, 3, 'cpu')
