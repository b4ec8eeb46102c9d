'''Constraint specifications for the CUDA engine.

Host-side mirror of safe_control_gym/envs/constraints.py: the same YAML `constraint_form`s and class
names, but a constraint here is a *description* -- rows of `g(z) = a.z - b <= 0` (or `|a.z| - b`, or the quadratic form `z'Qz - b`) over z = [state; input]
-- that is uploaded once (scg_set_constraints) and evaluated every step inside the kernels with the
reference's numerics: A, b cast to float32 (constraints.py:267-268), values rounded to 8 decimals
(constraints.py:109), violated when `g > 0` (`>=` if strict, constraints.py:126-129), input constraints
reading `current_noisy_physical_action` (constraints.py:168-169).
'''

from enum import Enum

import numpy as np

from safe_control_gym_b200 import _abi


class ConstrainedVariableType(str, Enum):
    STATE = 'state'
    INPUT = 'input'
    INPUT_AND_STATE = 'input_and_state'


class Constraint:
    '''Base: bookkeeping shared by all forms (constraints.py:34-83).'''

    def __init__(self, env, constrained_variable, strict=False, active_dims=None, tolerance=None, decimals=8, **kwargs):
        self.constrained_variable = ConstrainedVariableType(constrained_variable)
        self._nx, self._nu = env.state_dim, env.action_dim
        if self.constrained_variable == ConstrainedVariableType.STATE:
            self.dim = env.state_dim
        elif self.constrained_variable == ConstrainedVariableType.INPUT:
            self.dim = env.action_dim
        else:
            self.dim = env.state_dim + env.action_dim
        if decimals != 8:
            raise NotImplementedError('the CUDA engine rounds constraint values to 8 decimals (the reference default)')
        self.strict = strict
        self.decimals = decimals
        if active_dims is not None:
            if isinstance(active_dims, int):
                active_dims = [active_dims]
            assert isinstance(active_dims, (list, np.ndarray)), '[ERROR] active_dims is not a list/array.'
            assert len(active_dims) <= self.dim, '[ERROR] more active_dim than constrainable self.dim'
            assert all(isinstance(n, (int, np.integer)) for n in active_dims), '[ERROR] non-integer active_dim.'
            assert all(n < self.dim for n in active_dims), '[ERROR] active_dim not stricly smaller than self.dim.'
            assert len(active_dims) == len(set(active_dims)), '[ERROR] duplicates in active_dim'
            self.constraint_filter = np.eye(self.dim)[list(active_dims)]
            self.dim = len(active_dims)
        else:
            self.constraint_filter = np.eye(self.dim)
        self.tolerance = np.array(tolerance, ndmin=1) if tolerance is not None else None

    def reset(self):
        return

    # -- engine view -------------------------------------------------------------------------
    def _z_offset(self):
        '''Index of this constraint's variable inside z = [state; input].'''
        return self._nx if self.constrained_variable == ConstrainedVariableType.INPUT else 0

    def rows(self):
        '''List of (kind, strict, {z_col: coeff}, b) rows.'''
        raise NotImplementedError

    # -- host helpers kept for API parity (not on the hot path) ---------------------------------
    def get_symbolic_model(self):
        return self.sym_func

    def get_env_constraint_var(self, env):
        if self.constrained_variable == ConstrainedVariableType.STATE:
            return env.state
        if self.constrained_variable == ConstrainedVariableType.INPUT:
            return env.current_noisy_physical_action
        return np.concatenate([np.atleast_1d(env.state), np.atleast_1d(env.current_noisy_physical_action)])

    def get_value(self, env):
        x = np.array(self.get_env_constraint_var(env), ndmin=1, dtype=float)
        return np.round(np.atleast_1d(np.squeeze(self.sym_func(x))), decimals=self.decimals)

    def is_violated(self, env, c_value=None):
        if c_value is None:
            c_value = self.get_value(env)
        return bool(np.any(np.greater_equal(c_value, 0.)) if self.strict else np.any(np.greater(c_value, 0.)))

    def is_almost_active(self, env, c_value=None):
        if self.tolerance is None:
            return False
        if c_value is None:
            c_value = self.get_value(env)
        return bool(np.any(np.greater(c_value + self.tolerance, 0.)))

    def check_tolerance_shape(self):
        if self.tolerance is not None and len(self.tolerance) != self.num_constraints:
            raise ValueError('[ERROR] the tolerance dimension does not match the number of constraints.')


class LinearConstraint(Constraint):
    '''A @ x <= b (constraints.py:234-283).'''

    def __init__(self, env, A, b, constrained_variable, strict=False, active_dims=None, tolerance=None, decimals=8):
        super().__init__(env, constrained_variable, strict=strict, active_dims=active_dims, tolerance=tolerance, decimals=decimals)
        A = np.asarray(A, dtype=np.float32).reshape(-1, self.dim)
        b = np.asarray(b, dtype=np.float32).reshape(-1)
        assert A.shape[1] == self.dim, '[ERROR] A has the wrong dimension!'
        assert b.shape[0] == A.shape[0], '[ERROR] Dimension 0 of b does not match A!'
        self.A, self.b = A, b
        self.num_constraints = A.shape[0]
        self.sym_func = lambda x: self.A @ self.constraint_filter @ x - self.b
        self.check_tolerance_shape()

    def rows(self):
        M = self.A.astype(np.float64) @ self.constraint_filter   # float32 values promoted, as numpy does
        off = self._z_offset()
        out = []
        for r in range(M.shape[0]):
            cols = {int(off + c): float(M[r, c]) for c in np.nonzero(M[r])[0]}
            out.append((_abi.SCG_CON_LINEAR, bool(self.strict), cols, float(self.b[r])))
        return out


class BoundedConstraint(LinearConstraint):
    '''lb <= x <= ub as [-I; I] x <= [-lb; ub] (constraints.py:286-329).'''

    def __init__(self, env, lower_bounds, upper_bounds, constrained_variable, strict=False, active_dims=None, tolerance=None,
                 decimals=8):
        self.lower_bounds = np.array(lower_bounds, ndmin=1)
        self.upper_bounds = np.array(upper_bounds, ndmin=1)
        dim = self.lower_bounds.shape[0]
        if active_dims is not None and isinstance(active_dims, list):
            assert self.lower_bounds.shape[0] == len(active_dims), '[Error] active_dims and lower_bounds must have the same dimension.'
        if active_dims is not None and isinstance(active_dims, int):
            assert self.lower_bounds.shape[0] == 1 and self.upper_bounds.shape[0] == 1, \
                '[Error] active_dims and bounds must have the same dimension.'
        A = np.vstack((-np.eye(dim), np.eye(dim)))
        b = np.hstack((-self.lower_bounds, self.upper_bounds))
        super().__init__(env, A, b, constrained_variable, strict=strict, active_dims=active_dims, tolerance=tolerance, decimals=decimals)


class DefaultConstraint(BoundedConstraint):
    '''Bounds from the env's state_space / physical action bounds (constraints.py:332-397).'''

    def __init__(self, env, constrained_variable, lower_bounds=None, upper_bounds=None, strict=False, tolerance=None, decimals=8):
        cv = ConstrainedVariableType(constrained_variable)
        if cv == ConstrainedVariableType.STATE:
            space = env.state_space if hasattr(env, 'state_space') else env.observation_space
            low, high = space.low, space.high
        elif cv == ConstrainedVariableType.INPUT:
            low = np.asarray(env.physical_action_bounds[0], dtype=np.float32)
            high = np.asarray(env.physical_action_bounds[1], dtype=np.float32)
        else:
            raise NotImplementedError('[ERROR] DefaultConstraint can only be of type STATE or INPUT')
        n = low.shape[0]
        if upper_bounds is None:
            upper_bounds = high
        else:
            upper_bounds = np.array(upper_bounds, ndmin=1)
            assert len(upper_bounds) == n, '[ERROR]: Upper bound must have length equal to space dimension.'
        if lower_bounds is None:
            lower_bounds = low
        else:
            lower_bounds = np.array(lower_bounds, ndmin=1)
            assert len(lower_bounds) == n, '[ERROR]: Lower bound must have length equal to space dimension.'
        super().__init__(env, lower_bounds=lower_bounds.astype(np.float64), upper_bounds=upper_bounds.astype(np.float64),
                         constrained_variable=cv, strict=strict, active_dims=None, tolerance=tolerance, decimals=decimals)


class SymmetricStateConstraint(BoundedConstraint):
    '''|F x| <= bound, cartpole + rl_reward only (constraints.py:400-455).'''

    def __init__(self, env, constrained_variable, bound, strict=False, active_dims=None, tolerance=None, decimals=8, **kwargs):
        assert bound is not None
        self.bound = np.array(bound, ndmin=1)
        super().__init__(env, lower_bounds=-self.bound, upper_bounds=self.bound, constrained_variable=constrained_variable,
                         strict=strict, active_dims=active_dims, tolerance=tolerance, decimals=decimals)
        assert env.NAME == 'cartpole', '[ERROR] SymmetricStateConstraint is meant for CartPole environments'
        assert env.COST == 'rl_reward', '[ERROR] SymmetricStateConstraint is meant for RL environments'
        self.num_constraints = self.bound.shape[0]

    def get_value(self, env):
        return np.round(np.abs(self.constraint_filter @ np.asarray(env.state, dtype=float)) - self.bound, decimals=self.decimals)

    def check_tolerance_shape(self):
        if self.tolerance is not None and len(self.tolerance) != len(self.bound):
            raise ValueError('[ERROR] the tolerance dimension does not match the number of constraints.')

    def rows(self):
        out = []
        for r in range(self.constraint_filter.shape[0]):
            cols = {int(c): float(self.constraint_filter[r, c]) for c in np.nonzero(self.constraint_filter[r])[0]}
            out.append((_abi.SCG_CON_ABS, bool(self.strict), cols, float(self.bound[r])))
        return out


class QuadraticConstraint(Constraint):
    '''x' P x <= b (constraints.py:181-231).  Engine view: ONE row of kind SCG_CON_QUAD over z = [state; input] whose non-zeros are the
    entries of Q = F' P F, each carrying its pair of z indices.'''

    def __init__(self, env, P, b, constrained_variable, strict=False, active_dims=None, tolerance=None, decimals=8):
        super().__init__(env, constrained_variable, strict=strict, active_dims=active_dims, tolerance=tolerance, decimals=decimals)
        P = np.array(P, ndmin=1)
        assert P.shape == (self.dim, self.dim), '[ERROR] P has the wrong dimension!'
        assert isinstance(b, float), '[ERROR] b is not a scalar!'
        self.P, self.b = P, b
        self.num_constraints = 1
        self.sym_func = lambda x: x.T @ self.constraint_filter.T @ self.P @ self.constraint_filter @ x - self.b
        self.check_tolerance_shape()

    def rows(self):
        Q = self.constraint_filter.T @ np.asarray(self.P, dtype=np.float64) @ self.constraint_filter
        off = self._z_offset()
        cols = {}
        for i, j in zip(*np.nonzero(Q)):
            cols[int(off + i) | (int(off + j) << 8)] = float(Q[i, j])
        return [(_abi.SCG_CON_QUAD, bool(self.strict), cols, float(self.b))]


class ConstraintList:
    '''Ordered collection (constraints.py:471-636).'''

    def __init__(self, constraints):
        self.constraints = constraints
        self.constraint_lengths = [con.num_constraints for con in constraints]
        self.constraint_indices = np.cumsum(self.constraint_lengths[:-1])
        self.num_constraints = sum(self.constraint_lengths)
        self.state_constraints = [c for c in constraints if c.constrained_variable == ConstrainedVariableType.STATE]
        self.num_state_constraints = sum(c.num_constraints for c in self.state_constraints)
        self.input_constraints = [c for c in constraints if c.constrained_variable == ConstrainedVariableType.INPUT]
        self.num_input_constraints = sum(c.num_constraints for c in self.input_constraints)
        self.input_state_constraints = [c for c in constraints if c.constrained_variable == ConstrainedVariableType.INPUT_AND_STATE]
        self.num_input_state_constraints = sum(c.num_constraints for c in self.input_state_constraints)

    def __len__(self):
        return len(self.constraints)

    def get_all_symbolic_models(self):
        return [c.get_symbolic_model() for c in self.constraints]

    def get_state_constraint_symbolic_models(self):
        return [c.get_symbolic_model() for c in self.state_constraints]

    def get_input_constraint_symbolic_models(self):
        return [c.get_symbolic_model() for c in self.input_constraints]

    def get_input_and_state_constraint_symbolic_models(self):
        return [c.get_symbolic_model() for c in self.input_state_constraints]

    def get_values(self, env, only_state=False):
        cons = self.state_constraints if only_state else self.constraints
        return np.concatenate([c.get_value(env) for c in cons])

    def get_violations(self, env, only_state=False):
        cons = self.state_constraints if only_state else self.constraints
        return [c.is_violated(env) for c in cons]

    def is_violated(self, env, c_value=None):
        if c_value is not None:
            splits = np.split(c_value, self.constraint_indices)
            return any(c.is_violated(env, c_value=s) for c, s in zip(self.constraints, splits))
        return any(c.is_violated(env) for c in self.constraints)

    def is_almost_active(self, env, c_value=None):
        if c_value is not None:
            splits = np.split(c_value, self.constraint_indices)
            return any(c.is_almost_active(env, c_value=s) for c, s in zip(self.constraints, splits))
        return any(c.is_almost_active(env) for c in self.constraints)

    # -- engine view -------------------------------------------------------------------------
    def csr(self):
        '''Flattens all rows to the CSR arrays of scg_set_constraints().'''
        kind, strict, ptr, col, val, b = [], [], [0], [], [], []
        for con in self.constraints:
            for (k, s, cols, bb) in con.rows():
                kind.append(k)
                strict.append(1 if s else 0)
                for c in sorted(cols):
                    col.append(c)
                    val.append(cols[c])
                ptr.append(len(col))
                b.append(bb)
        if len(kind) > _abi.SCG_MAX_CON_ROWS or len(col) > _abi.SCG_MAX_CON_NNZ:
            raise ValueError('too many constraint rows for the CUDA engine')
        return (np.asarray(kind, np.int32), np.asarray(strict, np.int32), np.asarray(ptr, np.int32),
                np.asarray(col, np.int32), np.asarray(val, np.float64), np.asarray(b, np.float64))


GENERAL_CONSTRAINTS = {
    'linear_constraint': LinearConstraint,
    'quadratic_constraint': QuadraticConstraint,
    'bounded_constraint': BoundedConstraint,
    'default_constraint': DefaultConstraint,
}


def create_constraint_list(constraint_specs, available_constraints, env):
    '''YAML constraint specs -> ConstraintList (constraints.py:647-665).'''
    constraint_list = []
    for constraint in constraint_specs:
        assert isinstance(constraint, dict), '[ERROR]: Each constraint must be specified as a dict.'
        assert 'constraint_form' in constraint.keys(), '[ERROR]: Each constraint must have a key \'constraint_form\''
        con_form = constraint['constraint_form']
        assert con_form in available_constraints, '[ERROR]. constraint not in list of available constraints'
        cfg = {key: constraint[key] for key in constraint if key != 'constraint_form'}
        constraint_list.append(available_constraints[con_form](env, **cfg))
    return ConstraintList(constraint_list)
