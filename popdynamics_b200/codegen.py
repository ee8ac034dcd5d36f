# -*- coding: utf-8 -*-
"""
CUDA C emission for a ``CompiledModel`` (popdynamics_b200/compiler.py).

The generated text is the *model header* that the hand-written kernels in
``popdynamics_b200/csrc/kernels/*.cuh`` include as ``pd_model.h`` when NVRTC specialises them:

* ``PD_C / PD_NFLOW / PD_NPARAM / PD_NPM / PD_NSU`` sizes, the flat flow table as
  ``PD_FOR_EACH_FLOW`` X-macro rows ``(row, from, to)`` so every compartment index is a literal
  and the state stays in registers;
* ``pd_derivative``  -- the traced ``calculate_vars`` DAG followed by ``calculate_flows``
  (/root/reference/basepop.py:560-598) as straight-line fp64 code, program order preserved;
* ``pd_event_mults`` -- the same DAG evaluated once per stochastic step; the per-row event rates
  of ``calculate_events`` (/root/reference/basepop.py:703-735) are spelled as expressions in the
  flow-table rows so that they are formed where they are used (fewer live registers).

Parameters that are uniform over the ensemble are read from the kernel-argument constant bank
(``U.p[k]``); per-member (swept) parameters are loaded once per thread into ``pm[j]``.
"""

import math

from .compiler import ENTRY, KIND_NAMES, _BOOL_OPS

_BIN = {'add': '+', 'sub': '-', 'mul': '*', 'div': '/', 'lt': '<', 'le': '<=', 'gt': '>',
        'ge': '>=', 'eq': '==', 'ne': '!=', 'and': '&&', 'or': '||'}


def _literal(x):
    if math.isnan(x):
        return '__longlong_as_double(0x7ff8000000000000LL)'
    if math.isinf(x):
        return ('' if x > 0 else '-') + '__longlong_as_double(0x7ff0000000000000LL)'
    return '%s /* %r */' % (float(x).hex(), x)


_LEAVES = ('comp', 'comp2', 'param', 'time', 'su', 'const')


def _node_expr(g, node, name):
    op = node[0]
    a = name[node[1]]
    b = name[node[2]] if node[2] >= 0 else None
    c = name[node[3]] if node[3] >= 0 else None
    if op == 'div':
        return 'PD_DIV(%s, %s)' % (a, b)
    if op == 'ne' and g.nodes[node[2]][0] == 'const' and g.const_values[node[2]] == 0.0:
        return 'PD_NE0(%s)' % a           # the `if c` of CPython's compensated sum()
    if op in _BIN:
        return '%s %s %s' % (a, _BIN[op], b)
    if op == 'neg':
        return '-%s' % a
    if op == 'fabs':
        return 'fabs(%s)' % a
    if op == 'floor':
        return 'floor(%s)' % a
    if op == 'exp':
        return 'PD_EXP(%s)' % a
    if op == 'not':
        return '!%s' % a
    if op == 'isfinite':
        return 'PD_ISFINITE(%s)' % a
    if op == 'select':
        return '%s ? %s : %s' % (a, b, c)
    raise ValueError('cannot emit op %r' % (op,))


_COSTLY = ('exp', 'div')     # what makes a select arm worth a branch


def _emit_dag(cm, lines, roots):
    """Evaluation of the DAG nodes ``roots`` depend on, in creation (= program) order; returns
    node id -> C expression name.

    Straight-line code, with one exception: a ``select`` whose arm owns an expensive subtree (an
    ``exp`` or a division used by nothing else) is emitted as ``if / else`` with that subtree
    inside the branch.  These are the path joins of scale-up closures traced on time
    (``compiler._trace_scalar_function``: `early(x) if x < 1990 else late(x)`, the `arg > 10`
    early return of the sigmoid): their conditions depend on time only, so the lanes of a warp
    nearly always agree and the untaken sigmoid is simply not evaluated.  Values are unchanged --
    only unused ones are skipped.  The cheap selects of the compensated ``sum`` stay ternaries."""
    g = cm.graph
    member_slot = {k: j for j, k in enumerate(cm.member_param_slots())}
    needed, stack = set(), list(roots)
    while stack:
        n = stack.pop()
        if n in needed:
            continue
        needed.add(n)
        if g.nodes[n][0] not in _LEAVES:
            stack.extend(x for x in g.nodes[n][1:] if x >= 0)
    order = [n for n in cm.order if n in needed]

    def operands(n):
        node = g.nodes[n]
        return [] if node[0] in _LEAVES else [x for x in node[1:] if x >= 0]

    users = {n: set() for n in order}
    for n in order:
        for x in operands(n):
            users[x].add(n)
    for n in roots:
        users[n].add('root')

    owner = {}                # node -> (select node, arm index) that emits it inside its branch
    lazy = {}                 # select node -> (nodes of arm 1 in order, nodes of arm 2 in order)
    for s_node in order:
        node = g.nodes[s_node]
        if node[0] != 'select' or node[2] == node[3] or node[1] in (node[2], node[3]):
            continue
        arms = []
        for arm in (node[2], node[3]):
            reach, stack = set(), [arm]
            while stack:
                m = stack.pop()
                if m in reach or g.nodes[m][0] in ('comp', 'comp2', 'param', 'time', 'su'):
                    continue
                reach.add(m)
                stack.extend(operands(m))
            own = set()
            for m in sorted(reach, reverse=True):           # users before operands
                allowed = own | ({s_node} if m == arm else set())
                if users[m] <= allowed:
                    own.add(m)
            arms.append(own)
        if arms[0] & arms[1]:
            continue
        if not any(g.nodes[m][0] in _COSTLY for own in arms for m in own):
            continue
        for k, own in enumerate(arms):
            for m in own:
                owner.setdefault(m, (s_node, k))            # inner selects claimed theirs first
        lazy[s_node] = tuple(sorted(own) for own in arms)

    name = {}

    def emit(n, where, indent):
        node = g.nodes[n]
        op = node[0]
        if op == 'comp':
            name[n] = 'y[%d]' % node[1]
        elif op == 'comp2':
            name[n] = 'y2[%d]' % node[1]
        elif op == 'param':
            k = node[1]
            name[n] = ('pm[%d]' % member_slot[k]) if k in member_slot else ('U.p[%d]' % k)
        elif op == 'time':
            name[n] = 't'
        elif op == 'su':
            name[n] = 'su[%d]' % node[1]
        elif op == 'const':
            name[n] = 'v%d' % n
            lines.append('%sconst double v%d = %s;' % (indent, n, _literal(g.const_values[n])))
        elif n in lazy:
            name[n] = 'v%d' % n
            lines.append('%sdouble v%d;' % (indent, n))
            for k, keyword in ((0, 'if (%s) {' % name[node[1]]), (1, '} else {')):
                lines.append(indent + keyword)
                for m in lazy[n][k]:
                    if owner.get(m) == (n, k):
                        emit(m, (n, k), indent + '  ')
                lines.append('%s  v%d = %s;' % (indent, n, name[node[2 + k]]))
            lines.append(indent + '}')
        else:
            ctype = 'bool' if op in _BOOL_OPS else 'double'
            name[n] = 'v%d' % n
            lines.append('%sconst %s v%d = %s;' % (indent, ctype, n, _node_expr(g, node, name)))

    for n in order:
        if n not in owner:
            emit(n, None, '  ')
    return name


def _emit_checks(nodes, name, lines, default=False):
    lines.append('  bool ok = true;')
    if default:
        # the default checks() -- every compartment >= 0 (basepop.py:1016-1025) -- which a
        # kernel may evaluate itself (pd_explicit.cuh fuses it with the clamp of the update)
        lines.append('#ifndef PD_SKIP_DEFAULT_CHECKS')
    for n in nodes:
        lines.append('  ok = ok && %s;' % name[n])
    if default:
        lines.append('#endif')


def _is_default_checks(cm):
    """True when the checks are exactly `compartment c >= 0.0` for c = 0 .. C-1, in order: the
    reference's own BaseModel.checks() on a float state."""
    g = cm.graph
    if len(cm.check_nodes) != cm.n_comp or cm.n_comp == 0:
        return False
    for c, n in enumerate(cm.check_nodes):
        node = g.nodes[n]
        if node[0] != 'ge' or g.nodes[node[1]] [0] != 'comp' or g.nodes[node[1]][1] != c:
            return False
        if g.nodes[node[2]][0] != 'const' or g.const_values[node[2]] != 0.0:
            return False
    return True


_SIGNATURE = ('(const double (&y)[PD_C], const PdUniform& U, const double (&pm)[PD_NPM_DIM], '
              'const double t, const double* __restrict__ su, ')


def emit_model_header(cm):
    C, S = cm.n_comp, cm.n_flow
    npm = len(cm.member_param_slots())
    out = []
    w = out.append
    w('// ---- generated by popdynamics_b200/codegen.py: do not edit ----')
    w('// method: %s   compartments: %s' % (cm.method, ', '.join(cm.labels)))
    # a kernel may supply its own forms of the traced division and exp (kernels/pd_math.cuh)
    w('#ifndef PD_DIV')
    w('#define PD_DIV(a, b) ((a) / (b))')
    w('#endif')
    w('#ifndef PD_EXP')
    w('#define PD_EXP(x) exp(x)')
    w('#endif')
    w('#define PD_C %d' % C)
    w('#define PD_NFLOW %d' % S)
    w('#define PD_NFLOW_DIM %d' % max(S, 1))
    w('#define PD_NPARAM %d' % cm.n_param)
    w('#define PD_NPARAM_DIM %d' % max(cm.n_param, 1))
    w('#define PD_NPM %d' % npm)
    w('#define PD_NPM_DIM %d' % max(npm, 1))
    w('#define PD_NSU %d' % cm.n_su)
    w('#define PD_USES_TIME %d' % int(cm.uses_time))
    w('#define PD_INT_STATE %d' % int(cm.integer_state))
    w('#define PD_NAIVE_SUM %d' % int(cm.naive_sum))
    w('#define PD_HAS_CHECKS %d' % int(bool(cm.check_nodes)))
    default_checks = _is_default_checks(cm)
    w('#define PD_DEFAULT_CHECKS %d' % int(default_checks))
    # a kernel may supply bit-test forms of these two (same truth values, kernels/pd_math.cuh)
    w('#ifndef PD_ISFINITE')
    w('#define PD_ISFINITE(x) pd_isfinite(x)')
    w('#endif')
    w('#ifndef PD_NE0')
    w('#define PD_NE0(x) ((x) != 0.0)')
    w('#endif')
    w('struct PdUniform { double p[PD_NPARAM_DIM]; };')
    w('__device__ __forceinline__ bool pd_isfinite(double x) '
      '{ return fabs(x) <= 1.7976931348623157e308; }')
    w('')
    # ---- rate expression of every row: leaves are spelled inline (uniform parameters come from
    # the constant bank), computed vars go through mult[] filled once per step by pd_event_mults
    g = cm.graph
    member_slot = {k: j for j, k in enumerate(cm.member_param_slots())}
    mult_index = {}
    def leaf_text(n):
        node = g.nodes[n]
        if node[0] == 'param':
            k = node[1]
            return ('pm[%d]' % member_slot[k]) if k in member_slot else ('U.p[%d]' % k)
        if node[0] == 'const':
            return '(%s)' % _literal(g.const_values[n])
        if node[0] == 'comp':
            return 'yd[%d]' % node[1]
        if node[0] == 'time':
            return 'time'
        return None
    rate_expr = []
    for e in range(S):
        n = cm.flow_node[e]
        text = leaf_text(n)
        if text is None:
            if n not in mult_index:
                mult_index[n] = len(mult_index)
            text = 'mult[%d]' % mult_index[n]
        if cm.flow_kind[e] != ENTRY:
            text = 'yd[%d] * %s' % (cm.flow_from[e], text)
        rate_expr.append('(%s)' % text)
    w('#define PD_NMULT %d' % len(mult_index))
    w('#define PD_NMULT_DIM %d' % max(len(mult_index), 1))
    w('// flow table rows in canonical order; `rate` is an expression over yd[] (start-of-step')
    w('// counts as doubles), mult[], U.p[], pm[]:')
    w('//   ENTRY(row, to, is_int, rate) TRANSFER(row, from, to, rate) DEATH(row, from, rate)')
    rows = []
    for e in range(S):
        kind, frm, to = cm.flow_kind[e], cm.flow_from[e], cm.flow_to[e]
        if kind == ENTRY:
            rows.append('  ENTRY(%d, %d, %d, %s) /* %s */' % (e, to, cm.flow_is_int[e],
                                                            rate_expr[e], cm.flow_names[e]))
        elif to >= 0:
            rows.append('  TRANSFER(%d, %d, %d, %s) /* %s: %s */'
                        % (e, frm, to, rate_expr[e], KIND_NAMES[kind], cm.flow_names[e]))
        else:
            rows.append('  DEATH(%d, %d, %s) /* %s: %s */' % (e, frm, rate_expr[e],
                                                            KIND_NAMES[kind], cm.flow_names[e]))
    w('#define PD_FOR_EACH_FLOW(ENTRY, TRANSFER, DEATH) \\')
    w(' \\\n'.join(rows) if rows else '  /* no flows */')
    w('')
    # ---- the rows the stochastic kernels walk: a transfer / death row whose rate multiplier is a
    # uniform parameter (or constant) equal to zero can never fire (rate > 0 is false for every
    # member and every step, basepop.py:712-733), so it is dropped and the rest renumbered densely
    # in the same order.  Entry rows always stay (basepop.py:706-707).
    def never_fires(e):
        if cm.flow_kind[e] == ENTRY:
            return False
        node = g.nodes[cm.flow_node[e]]
        if node[0] == 'param' and node[1] not in member_slot:
            return float(cm.uniform_params()[node[1]]) == 0.0
        if node[0] == 'const':
            return float(g.const_values[cm.flow_node[e]]) == 0.0
        return False
    # A transfer / death row whose rate is count * (uniform parameter) has a Poisson code that
    # depends on the integer count only (for a constant step): the tau-leap kernel may look it up
    # in a table built once per launch instead of evaluating exp every step.  PD_ROW_TAB_<row> is
    # the table of that row (one table per distinct parameter slot), -1 for the other rows.
    # An entry row whose rate is a uniform parameter / constant has ONE code per dt for the whole
    # ensemble: PD_ROW_UNI_<row> is its slot behind the tables, -1 for the other rows.
    # A transfer / death row whose rate is count * (count' * uniform parameter) -- a force of
    # infection -- has a code that depends on two integer counts: PD_ROW_TAB2_<row> is its pair
    # table (one per distinct (compartment, parameter) inside the rate), PD_ROW_TAB2B_<row> the
    # compartment inside the rate, -1 for the other rows.
    tau_rows, tab_params, row_tab, uni_rates, row_uni = [], [], [], [], []
    tab2_keys, row_tab2 = [], []

    def pair_of(node):
        """(compartment, uniform parameter slot) when `node` is comp * param or param * comp."""
        if node[0] != 'mul':
            return None
        x, y = g.nodes[node[1]], g.nodes[node[2]]
        for c_node, p_node in ((x, y), (y, x)):
            if c_node[0] == 'comp' and p_node[0] == 'param' and p_node[1] not in member_slot:
                return (c_node[1], p_node[1])
        return None
    for e in range(S):
        if never_fires(e):
            continue
        kind, frm, to = cm.flow_kind[e], cm.flow_from[e], cm.flow_to[e]
        node = g.nodes[cm.flow_node[e]]
        uniform_leaf = (node[0] == 'param' and node[1] not in member_slot) or node[0] == 'const'
        if kind != ENTRY and node[0] == 'param' and node[1] not in member_slot:
            if node[1] not in tab_params:
                tab_params.append(node[1])
            row_tab.append(tab_params.index(node[1]))
        else:
            row_tab.append(-1)
        pair = pair_of(node) if (kind != ENTRY and cm.integer_state) else None
        if pair is not None:
            if pair not in tab2_keys:
                tab2_keys.append(pair)
            row_tab2.append((tab2_keys.index(pair), pair[0]))
        else:
            row_tab2.append((-1, -1))
        if kind == ENTRY and uniform_leaf:
            row_uni.append(len(uni_rates))
            uni_rates.append(rate_expr[e])
        else:
            row_uni.append(-1)
        r = len(tau_rows)
        if kind == ENTRY:
            tau_rows.append('  ENTRY(%d, %d, %d, %s) /* %s */' % (r, to, cm.flow_is_int[e],
                                                                rate_expr[e], cm.flow_names[e]))
            frm = -1
        elif to >= 0:
            tau_rows.append('  TRANSFER(%d, %d, %d, %s) /* %s */' % (r, frm, to, rate_expr[e],
                                                                   cm.flow_names[e]))
        else:
            tau_rows.append('  DEATH(%d, %d, %s) /* %s */' % (r, frm, rate_expr[e],
                                                            cm.flow_names[e]))
    w('#define PD_LIVE_NROW %d' % len(tau_rows))
    w('#define PD_LIVE_NROW_DIM %d' % max(len(tau_rows), 1))
    w('#define PD_NTAB %d' % len(tab_params))
    w('#define PD_NTAB_DIM %d' % max(len(tab_params), 1))
    w('#define PD_TAB_PARAMS { %s }' % (', '.join(str(k) for k in tab_params) or '0'))
    for r, t in enumerate(row_tab):
        w('#define PD_ROW_TAB_%d %d' % (r, t))
    w('#define PD_NTAB2 %d' % len(tab2_keys))
    w('#define PD_NTAB2_DIM %d' % max(len(tab2_keys), 1))
    w('#define PD_TAB2_COMPS { %s }' % (', '.join(str(c) for c, _ in tab2_keys) or '0'))
    w('#define PD_TAB2_PARAMS { %s }' % (', '.join(str(k) for _, k in tab2_keys) or '0'))
    for r, (t, b) in enumerate(row_tab2):
        w('#define PD_ROW_TAB2_%d %d' % (r, t))
        w('#define PD_ROW_TAB2B_%d %d' % (r, b))
    w('#define PD_NUNI %d' % len(uni_rates))
    w('#define PD_NUNI_DIM %d' % max(len(uni_rates), 1))
    w('#define PD_UNI_RATES { %s }' % (', '.join(uni_rates) or '0.0'))
    for r, j in enumerate(row_uni):
        w('#define PD_ROW_UNI_%d %d' % (r, j))
    w('#define PD_FOR_EACH_LIVE_ROW(ENTRY, TRANSFER, DEATH) \\')
    w(' \\\n'.join(tau_rows) if tau_rows else '  /* no rows */')
    w('')

    # ---- derivative (explicit Euler) -----------------------------------------------------
    w('__device__ __forceinline__ bool pd_derivative' + _SIGNATURE + 'double (&f)[PD_C]) {')
    body = []
    name = _emit_dag(cm, body, list(cm.flow_node) + list(cm.check_nodes))
    _emit_checks(cm.check_nodes, name, body, default=default_checks)
    touched = [False] * C
    body.append('  double val;')

    def plus(i, expr):
        if touched[i]:
            body.append('  f[%d] += %s;' % (i, expr))
        else:
            body.append('  f[%d] = %s;' % (i, expr))
            touched[i] = True

    def minus(i, expr):
        if touched[i]:
            body.append('  f[%d] -= %s;' % (i, expr))
        else:
            body.append('  f[%d] = -%s;' % (i, expr))
            touched[i] = True

    for e in range(S):
        kind, frm, to = cm.flow_kind[e], cm.flow_from[e], cm.flow_to[e]
        rate = name[cm.flow_node[e]]
        body.append('  // %s' % cm.flow_names[e])
        if kind == ENTRY:
            plus(to, rate)
        else:
            body.append('  val = y[%d] * %s;' % (frm, rate))
            minus(frm, 'val')
            if to >= 0:
                plus(to, 'val')
    for i in range(C):
        if not touched[i]:
            body.append('  f[%d] = 0.0;' % i)
    body.append('  return ok;')
    out.extend(body)
    w('}')
    w('')

    # ---- diagnostics pass (basepop.py:911-960): every var and the net flow per compartment ---
    w('#define PD_NDIAG %d' % len(cm.diag_nodes))
    w('#define PD_NDIAG_DIM %d' % max(len(cm.diag_nodes), 1))
    if cm.diagnostics:
        w('__device__ __forceinline__ void pd_diag' + _SIGNATURE +
          'double (&out)[PD_NDIAG_DIM], double (&flow)[PD_C]) {')
        body = []
        name = _emit_dag(cm, body, list(cm.diag_nodes) + list(cm.diag_flow_nodes))
        for k, n in enumerate(cm.diag_nodes):
            body.append('  out[%d] = %s;  // %s' % (k, name[n], cm.diag_labels[k]))
        for c, n in enumerate(cm.diag_flow_nodes):
            body.append('  flow[%d] = %s;' % (c, name[n]))
        out.extend(body)
        w('}')
        w('')

    # ---- diagnostics vars of the FINAL state, written by the ODE kernels in PD_OUT_FINAL_VARS ---
    w('#define PD_NFVAR %d' % len(cm.final_var_nodes))
    w('#define PD_NFVAR_DIM %d' % max(len(cm.final_var_nodes), 1))
    if cm.final_var_nodes:
        w('__device__ __forceinline__ void pd_final_vars' + _SIGNATURE +
          'double (&out)[PD_NFVAR_DIM]) {')
        body = []
        name = _emit_dag(cm, body, list(cm.final_var_nodes))
        for k, n in enumerate(cm.final_var_nodes):
            body.append('  out[%d] = %s;  // %s' % (k, name[n], cm.final_var_labels[k]))
        out.extend(body)
        w('}')
        w('')

    # ---- per-step multipliers of the stochastic integrators ------------------------------
    w('__device__ __forceinline__ bool pd_event_mults' + _SIGNATURE +
      'double (&mult)[PD_NMULT_DIM]) {')
    body = []
    name = _emit_dag(cm, body, list(mult_index) + list(cm.check_nodes))
    _emit_checks(cm.check_nodes, name, body)
    for n, k in sorted(mult_index.items(), key=lambda kv: kv[1]):
        body.append('  mult[%d] = %s;' % (k, name[n]))
    body.append('  return ok;')
    out.extend(body)
    w('}')
    w('')

    # ---- a user-defined checks() in the stochastic loops (basepop.py:789, :846): evaluated after
    # the update on the updated counts y2; vars are those of the step's start (from y)
    w('#define PD_HAS_STEP_CHECKS %d' % int(bool(cm.post_check_nodes)))
    if cm.post_check_nodes:
        w('__device__ __forceinline__ bool pd_step_checks(const double (&y)[PD_C], '
          'const double (&y2)[PD_C], const PdUniform& U, const double (&pm)[PD_NPM_DIM], '
          'const double t, const double* __restrict__ su) {')
        body = []
        name = _emit_dag(cm, body, list(cm.post_check_nodes))
        _emit_checks(cm.post_check_nodes, name, body)
        body.append('  return ok;')
        out.extend(body)
        w('}')
        w('')
    return '\n'.join(out) + '\n'
