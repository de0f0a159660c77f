"""Record event-by-event traces from the RUNNING reference (JeLLyFysh) -- golden-vector generator.

Runs only where the reference checkout is available (this container): `REF` must point to a copy of the
reference whose three cffi extensions are built (baseline/_ref, see tests/golden/make_golden.py). Nothing in
tests/, bench.py or smoke() imports this at run time on the GPU box; only its committed outputs
(tests/golden/*.npz) travel.

How the reference is made reproducible: the reference draws from the global `random` module in
an order that depends on set iteration. The recorder replaces the module-level functions of `random` by a
counter-based stream keyed by WHAT the draw is for, not by WHEN it happens:

    u = Philox4x32-10(key=(stream, seed), counter=(event_lo, event_hi, slot, block))

with slot = (kind << 24 | index) as in include/ecmc.h and `event` = number of committed device events of the
chain. The handlers' send_event_time / send_out_state are wrapped only to set that context; every arithmetic
step is the unmodified reference.
"""
import configparser
import contextlib
import io
import os
import random
import sys

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Pure-Python Philox4x32-10 (Salmon et al., SC'11); independent of the C oracle's implementation."""
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def stream_word(seed, stream, event, slot, index):
    c = philox4x32_10((event & MASK, (event >> 32) & MASK, slot, index >> 2), (stream, seed))
    return c[index & 3]


def stream_double(seed, stream, event, slot, index):
    c = philox4x32_10((event & MASK, (event >> 32) & MASK, slot, index >> 1), (stream, seed))
    a, b = c[2 * (index & 1)] >> 5, c[2 * (index & 1) + 1] >> 6
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)


SLOT_PAIR_TIME, SLOT_VETO_TIME, SLOT_VETO_CHOICE, SLOT_CONFIRM, SLOT_END_OF_CHAIN, SLOT_LIFTING = 1, 2, 3, 4, 5, 6
SLOT_FACTOR_TIME, SLOT_BENDING_TIME = 7, 8
SLOT_SWITCH = 9  # words -> random.choice over the leaves when the root unit hands over to one of them
SLOT_INIT = 15  # initial random positions (host side only)


def make_slot(kind, index=0):
    return ((kind << 24) | (index & 0xFFFFFF)) & MASK


class SlotRandom(random.Random):
    """random.Random whose primitives read the slot-keyed stream. expovariate / uniform / choice / randint
    are CPython's own methods on top of random() and getrandbits()."""

    def __init__(self, seed, stream):
        self.seed_word, self.stream = seed, stream
        self.event = 0
        self.dslot = self.wslot = None
        self.di = self.wi = 0
        self.draw_log = []
        super().__init__(0)

    def set_context(self, event, dslot, wslot=None):
        self.event, self.dslot, self.wslot, self.di, self.wi = event, dslot, wslot, 0, 0

    def clear_context(self):
        self.dslot = self.wslot = None

    def random(self):
        if self.dslot is None:
            raise RuntimeError("reference drew a double outside of an instrumented context")
        u = stream_double(self.seed_word, self.stream, self.event, self.dslot, self.di)
        self.di += 1
        return u

    def getrandbits(self, k):
        if self.wslot is None:
            raise RuntimeError("reference drew bits outside of an instrumented context")
        assert 0 < k <= 32
        w = stream_word(self.seed_word, self.stream, self.event, self.wslot, self.wi)
        self.wi += 1
        return w >> (32 - k)

    def seed(self, *args, **kwargs):  # called by Random.__init__
        return None


EVENT_PAIR, EVENT_CELL_VETO, EVENT_CELL_BOUNDARY, EVENT_END_OF_CHAIN, EVENT_CELL_BOUNDING = 1, 2, 3, 4, 5
EVENT_BOND, EVENT_FACTOR_PAIR, EVENT_BENDING = 6, 7, 8
EVENT_SWITCH = 9  # RootLeafUnitActiveSwitcher: the root unit / one leaf unit of the same object takes over
HOST_EVENT = 0

RECORD_DTYPE = np.dtype([("kind", "<i4"), ("target", "<i4"), ("target_cell", "<i4"), ("accepted", "<i4"),
                         ("n_candidates", "<i4"), ("new_active", "<i4"), ("new_direction", "<i4"),
                         ("reserved", "<i4"), ("time_q", "<f8"), ("time_r", "<f8"), ("active_pos", "<f8", (3,))])


def import_reference(ref_root):
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import warnings
    warnings.filterwarnings("ignore")
    import jellyfysh  # noqa: F401
    return jellyfysh


class ReferenceRun:
    """Build the reference object graph from INI text and run it with full instrumentation."""

    def __init__(self, ref_root, ini_text, seed=0, stream=0, positions=None, composites=None):
        """positions: start configuration of point masses, fed through setting.random_position. composites: start
        configuration of composite point objects as (roots[n_roots][D], leaves[n_roots][nodes_per_root][D]), fed
        through the fill_root_node method of the INI's random node creator (the only patched reference method: it
        replaces the random draw of a molecule by the given one, charges as the creator assigns them)."""
        import_reference(ref_root)
        from jellyfysh.base import factory
        from jellyfysh.base.strings import to_camel_case
        import jellyfysh.setting as setting
        from jellyfysh.activator.tagger.factor_type_maps import FactorTypeMaps
        self.setting = setting
        setting.reset()
        FactorTypeMaps._instance = None
        factory.used_sections.clear()
        self.rng = SlotRandom(seed, stream)
        self._patch_random()
        config = configparser.ConfigParser()
        config.read_string(ini_text)
        self.config = config
        factory.build_from_config(config, to_camel_case(config.get("Run", "setting")), "jellyfysh.setting")
        if positions is not None:
            pos_iter = iter([list(map(float, p)) for p in positions])
            setting.random_position = lambda: next(pos_iter)
        self._unpatch_creators = []
        if composites is not None:
            self._patch_node_creators(composites)
        # initial positions of RandomInputHandler come from setting.random_position() -> random.uniform
        self.rng.set_context(0, make_slot(SLOT_INIT))
        with contextlib.redirect_stdout(io.StringIO()):
            self.mediator = factory.build_from_config(config, to_camel_case(config.get("Run", "mediator")),
                                                      "jellyfysh.mediator")
        self.rng.clear_context()
        self.events = 0  # committed device events
        self.records = []
        self.iterations = []  # (winner class name, event time float)
        self.host_times = []  # (committed device events so far, quotient, remainder) of sampling events
        self._instrument()

    def _patch_node_creators(self, composites):
        import configs
        self._unpatch_creators.append(configs.patch_composite_start(composites))

    def leaf_id(self, identifier):
        """Flat leaf identifier root * nodes_per_root + child (the root identifier for point masses)."""
        if len(identifier) == 1:
            return identifier[0]
        return identifier[0] * self.setting.number_of_nodes_per_root_node + identifier[1]

    # -- random -------------------------------------------------------------------------------------
    def _patch_random(self):
        import jellyfysh.event_handler.single_independent_active_periodic_direction_end_of_chain_event_handler as eoc
        self._saved = {name: getattr(random, name) for name in
                       ("random", "uniform", "expovariate", "choice", "randint", "getrandbits")}
        self._saved_eoc_randint = eoc.randint
        for name in self._saved:
            setattr(random, name, getattr(self.rng, name))
        eoc.randint = self.rng.randint
        self._eoc_module = eoc

    def close(self):
        for name, fn in self._saved.items():
            setattr(random, name, fn)
        self._eoc_module.randint = self._saved_eoc_randint
        for restore in self._unpatch_creators:
            restore()
        self.setting.reset()

    # -- instrumentation ------------------------------------------------------------------------------
    def _kind_of(self, handler):
        names = {cls.__name__ for cls in type(handler).__mro__}
        if "CellVetoEventHandler" in names:
            return EVENT_CELL_VETO
        if names & {"TwoLeafUnitCellBoundingPotentialEventHandler", "TwoCompositeObjectCellBoundingPotentialEventHandler"}:
            return EVENT_CELL_BOUNDING
        if "CellBoundaryEventHandler" in names:
            return EVENT_CELL_BOUNDARY
        if "EndOfChainEventHandler" in names:
            return EVENT_END_OF_CHAIN
        if "FixedSeparationsEventHandlerWithPiecewiseConstantBoundingPotential" in names:
            return EVENT_BENDING
        if names & {"TwoCompositeObjectSummedBoundingPotentialEventHandler",
                    "RootUnitActiveTwoCompositeObjectSummedBoundingPotentialEventHandler"}:
            return EVENT_PAIR
        if "RootLeafUnitActiveSwitcher" in names:
            return EVENT_SWITCH
        if "TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential" in names:
            # a two-leaf factor between objects whose candidate comes from a piecewise constant bound (the Lennard-Jones
            # factor of the oxygens in nearby cells, water/coulomb_power_bounded_lj_cell_bounded.ini)
            return EVENT_FACTOR_PAIR
        if names & {"TwoLeafUnitEventHandler", "TwoLeafUnitBoundingPotentialEventHandler"}:
            local = self._factor_map_handlers().get(id(handler))
            if self.setting.number_of_node_levels == 1:
                # point masses without a cell system: the factor type map lists the pair factors themselves
                # (coulomb_atoms/power_bounded.ini, "[0, 1], Coulomb")
                return EVENT_PAIR
            if local is False and "TwoLeafUnitBoundingPotentialEventHandler" in names:
                # bounded leaf-to-leaf factors between different objects (dipoles/atom_factors.ini): pair events
                return EVENT_PAIR
            return EVENT_PAIR if local is None else (EVENT_BOND if local else EVENT_FACTOR_PAIR)
        return HOST_EVENT

    def _factor_map_handlers(self):
        """id -> is the factor local (intramolecular)? for the handlers of every FactorTypeMapInStateTagger."""
        if getattr(self, "_factor_ids", None) is None:
            self._factor_ids = {}
            for tagger in self.mediator._activator._taggers:
                if "FactorTypeMapInStateTagger" in {cls.__name__ for cls in type(tagger).__mro__}:
                    local = bool(getattr(tagger._factor_type_map, "_local", False))
                    self._factor_ids.update({id(h): local for h in tagger.get_event_handlers()})
        return self._factor_ids

    def _target_root_of(self, in_state):
        """The root of the composite object in the in-state that holds no active leaf unit."""
        from jellyfysh.base.node import yield_leaf_nodes
        # an in-state from the cell taggers holds one branch per object, one from a factor type map one branch per leaf
        active_roots = {cnode.value.identifier[0] for cnode in in_state
                        if any(leaf.value.velocity is not None for leaf in yield_leaf_nodes(cnode))}
        for cnode in in_state:
            if cnode.value.identifier[0] not in active_roots:
                return cnode.value.identifier[0]
        raise RuntimeError("composite in-state without target")

    @staticmethod
    def _is_composite_pair(handler):
        return bool({"TwoCompositeObjectSummedBoundingPotentialEventHandler",
                     "RootUnitActiveTwoCompositeObjectSummedBoundingPotentialEventHandler"}
                    & {c.__name__ for c in type(handler).__mro__})

    @staticmethod
    def _is_composite_cell_bounding(handler):
        return "TwoCompositeObjectCellBoundingPotentialEventHandler" in {c.__name__ for c in type(handler).__mro__}

    def _is_leaf_pair_between_objects(self, handler):
        names = {c.__name__ for c in type(handler).__mro__}
        return (self.setting.number_of_node_levels == 2 and "TwoLeafUnitBoundingPotentialEventHandler" in names
                and self._factor_map_handlers().get(id(handler)) is False)

    def _target_of_pair(self, in_state):
        from jellyfysh.base.node import yield_leaf_nodes
        for cnode in in_state:
            for leaf in yield_leaf_nodes(cnode):
                if leaf.value.velocity is None:
                    return self.leaf_id(leaf.value.identifier)
        raise RuntimeError("pair in-state without target")

    def _moving_child_of_pair(self, in_state):
        from jellyfysh.base.node import yield_leaf_nodes
        for cnode in in_state:
            for leaf in yield_leaf_nodes(cnode):
                if leaf.value.velocity is not None:
                    return leaf.value.identifier[-1]
        raise RuntimeError("pair in-state without moving leaf")

    def _instrument(self):
        med = self.mediator
        run = self
        handlers = med._activator.get_event_handlers()
        self._pushed = []
        for h in handlers:
            kind = self._kind_of(h)
            orig_time, orig_out = h.send_event_time, h.send_out_state

            def send_event_time(*args, _h=h, _kind=kind, _orig=orig_time):
                if (_kind == EVENT_PAIR and run._is_composite_pair(_h)) or \
                        (_kind == EVENT_CELL_BOUNDING and run._is_composite_cell_bounding(_h)):
                    run.rng.set_context(run.events, make_slot(SLOT_PAIR_TIME, run._target_root_of(args[0])))
                elif _kind == EVENT_PAIR and run._is_leaf_pair_between_objects(_h):
                    # keyed like the composite-object handler: (pair time, target object), double = target child
                    leaf, npr = run._target_of_pair(args[0]), run.setting.number_of_nodes_per_root_node
                    run.rng.set_context(run.events, make_slot(SLOT_PAIR_TIME, leaf // npr))
                    run.rng.di = leaf % npr
                elif _kind == EVENT_CELL_BOUNDING and run.setting.number_of_node_levels == 2:
                    # a leaf-level cell-bounding handler next to composite-object pair handlers (which use the pair-time
                    # slots of the objects): double 1 of the factor-time slot of the target leaf
                    run.rng.set_context(run.events, make_slot(SLOT_FACTOR_TIME, run._target_of_pair(args[0])))
                    run.rng.di = 1
                elif _kind in (EVENT_PAIR, EVENT_CELL_BOUNDING):
                    run.rng.set_context(run.events, make_slot(SLOT_PAIR_TIME, run._target_of_pair(args[0])))
                elif _kind in (EVENT_BOND, EVENT_FACTOR_PAIR):
                    run.rng.set_context(run.events, make_slot(SLOT_FACTOR_TIME, run._target_of_pair(args[0])))
                    if "RootUnitActiveTwoLeafUnitEventHandler" in {c.__name__ for c in type(_h).__mro__}:
                        # the root unit is active: several moving leaves may meet the same target leaf, the double is
                        # the child index of the moving leaf of this factor
                        run.rng.di = run._moving_child_of_pair(args[0])
                elif _kind == EVENT_BENDING:
                    run.rng.set_context(run.events, make_slot(SLOT_BENDING_TIME))
                elif _kind == EVENT_CELL_VETO:
                    run.rng.set_context(run.events, make_slot(SLOT_VETO_TIME), make_slot(SLOT_VETO_CHOICE))
                elif _kind == EVENT_END_OF_CHAIN:
                    run.rng.set_context(run.events, None, make_slot(SLOT_END_OF_CHAIN))
                else:
                    run.rng.clear_context()
                try:
                    return _orig(*args)
                finally:
                    run.rng.clear_context()

            def send_out_state(*args, _h=h, _kind=kind, _orig=orig_out):
                if _kind in (EVENT_PAIR, EVENT_CELL_VETO, EVENT_CELL_BOUNDING, EVENT_BENDING) or \
                        "TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential" in {c.__name__ for c in type(_h).__mro__}:
                    # confirmation, then the draws of the lifting scheme, in call order
                    run.rng.set_context(run.events, make_slot(SLOT_CONFIRM))
                elif _kind == EVENT_SWITCH:
                    run.rng.set_context(run.events, None, make_slot(SLOT_SWITCH))
                else:
                    run.rng.clear_context()
                try:
                    return _orig(*args)
                finally:
                    run.rng.clear_context()

            h.send_event_time = send_event_time
            h.send_out_state = send_out_state

        sched = med._scheduler
        orig_push, orig_get = sched.push_event, sched.get_succeeding_event

        def push_event(time, handler):
            run._pushed.append((time, handler))
            return orig_push(time, handler)

        def get_succeeding_event():
            winner = orig_get()
            run._on_winner(winner)
            return winner

        sched.push_event = push_event
        sched.get_succeeding_event = get_succeeding_event

        sh = med._state_handler
        orig_insert = sh.insert_into_global_state

        def insert_into_global_state(out_state):
            # the reference method recurses into children through self.insert_into_global_state
            run._insert_depth += 1
            try:
                orig_insert(out_state)
            finally:
                run._insert_depth -= 1
            if run._insert_depth == 0 and run._current is not None:
                run._on_commit()

        self._insert_depth = 0
        self._current = None
        sh.insert_into_global_state = insert_into_global_state

    def _active(self):
        sh = self.mediator._state_handler
        ids = sorted(sh._lifting_state._lifting_dictionary.keys(), key=lambda i: (len(i), i))
        leaves = [i for i in ids if len(i) == self.setting.number_of_node_levels]
        # the active leaf and its ancestors -- or, with the root unit active, the root and all of its leaves (the first
        # leaf then stands for the object)
        self.mode = int(len(leaves) > 1)
        assert len(ids) == self.setting.number_of_node_levels or \
            len(leaves) == self.setting.number_of_nodes_per_root_node
        leaf = leaves[0]
        velocity, stamp = sh._lifting_state.get(leaf)
        pos = sh._physical_state.get(leaf).value.position
        direction = [i for i, v in enumerate(velocity) if v != 0.0][0]
        return self.leaf_id(leaf), direction, list(pos), stamp

    def _cell_index(self, cell):
        cells = self._cells()
        return sum(cell.identifier[d] * cells._cumulative_product[d] for d in range(self.setting.dimension))

    def _cells(self):
        return self.mediator._activator._internal_states[0].cells

    def _on_winner(self, winner):
        from math import isinf
        kind = self._kind_of(winner)
        pushed = [(t, h) for t, h in self._pushed if not isinf(t.quotient)]
        self._pushed = []
        name = type(winner).__name__
        if kind == HOST_EVENT:
            self._current = None
            self.iterations.append((name, None))
            t = getattr(winner, "_event_time", None)
            if t is not None and "Sampling" in name:
                self.host_times.append((self.events, t.quotient, t.remainder))
            return
        n_interaction = sum(1 for _, h in pushed
                            if self._kind_of(h) not in (HOST_EVENT, EVENT_END_OF_CHAIN, EVENT_SWITCH))
        rec = np.zeros((), dtype=RECORD_DTYPE)
        rec["kind"] = kind
        rec["target"] = -1
        rec["target_cell"] = -1
        # + the end of chain (+ the switcher of a program with a root-unit-active mode): persistent candidates
        rec["n_candidates"] = n_interaction + 1 + int(self._has_switcher())
        rec["time_q"] = winner._event_time.quotient
        rec["time_r"] = winner._event_time.remainder
        old_active = self._active()[0] if kind != EVENT_END_OF_CHAIN or self.events >= 0 else -1
        if kind == EVENT_PAIR and self._is_composite_pair(winner):
            rec["target"] = winner._target_leaf_units[0].identifier[0]  # the target composite object (root)
        elif kind in (EVENT_PAIR, EVENT_BOND, EVENT_FACTOR_PAIR):
            rec["target"] = [self.leaf_id(u.identifier) for u in winner._leaf_units if u.velocity is None][0]
        elif kind == EVENT_CELL_BOUNDING and self._is_composite_cell_bounding(winner):
            rec["target"] = winner._target_leaf_units[0].identifier[0]  # the target composite object (root)
            rec["target_cell"] = self._cell_index(winner._relative_cell)
        elif kind == EVENT_CELL_BOUNDING:
            rec["target"] = [u.identifier[0] for u in winner._leaf_units if u.velocity is None][0]
            rec["target_cell"] = self._cell_index(winner._relative_cell)
        elif kind == EVENT_CELL_VETO:
            cell = self.mediator._out_state_arguments[winner][0]
            rec["target_cell"] = self._cell_index(cell)
            occ = self.mediator._activator.get_info_internal_state(winner, cell)
            rec["target"] = occ[0][0] if occ else -1
        elif kind == EVENT_END_OF_CHAIN:
            identifier = self.mediator._out_state_arguments[winner][0][0]
            if len(identifier) < self.setting.number_of_node_levels:  # a root unit takes over: its first leaf
                identifier = tuple(identifier) + (0,)
            rec["target"] = self.leaf_id(identifier)
        self._current = (rec, old_active, kind)
        self.iterations.append((name, float(rec["time_q"] + rec["time_r"])))

    def _on_commit(self):
        rec, old_active, kind = self._current
        self._current = None
        new_active, direction, _, _ = self._active()
        sh = self.mediator._state_handler
        pos = sh._physical_state.get(self._identifier_of(old_active)).value.position
        rec["new_active"] = new_active
        rec["new_direction"] = direction
        rec["accepted"] = int(new_active != old_active or kind in (EVENT_END_OF_CHAIN, EVENT_SWITCH))
        rec["reserved"] = self.mode  # 1: the root unit is active after the event
        if kind == EVENT_CELL_BOUNDARY:
            cells = self._cells()
            cell_level = self.mediator._activator._internal_states[0].cell_level
            on_cell_level = sh._physical_state.get(self._identifier_of(old_active)[:cell_level]).value.position
            rec["target_cell"] = self._cell_index(cells.position_to_cell(on_cell_level))
        for d in range(self.setting.dimension):
            rec["active_pos"][d] = pos[d]
        self.records.append(rec.copy())
        self.events += 1

    def _has_switcher(self):
        if getattr(self, "_switcher", None) is None:
            self._switcher = any(self._kind_of(h) == EVENT_SWITCH for h in self.mediator._activator.get_event_handlers())
        return self._switcher

    def _identifier_of(self, leaf):
        if self.setting.number_of_node_levels == 1:
            return (leaf,)
        npr = self.setting.number_of_nodes_per_root_node
        return (leaf // npr, leaf % npr)

    # -- state snapshots -------------------------------------------------------------------------------
    def positions(self):
        """Leaf positions in flat leaf order."""
        sh = self.mediator._state_handler
        n = self.setting.number_of_root_nodes
        if self.setting.number_of_node_levels == 1:
            return np.array([sh._physical_state.get((i,)).value.position for i in range(n)], dtype=np.float64)
        npr = self.setting.number_of_nodes_per_root_node
        return np.array([sh._physical_state.get((i, k)).value.position for i in range(n) for k in range(npr)],
                        dtype=np.float64)

    def roots(self):
        sh = self.mediator._state_handler
        n = self.setting.number_of_root_nodes
        return np.array([sh._physical_state.get((i,)).value.position for i in range(n)], dtype=np.float64)

    def charges(self, name):
        sh = self.mediator._state_handler
        n = self.setting.number_of_root_nodes
        return np.array([sh._physical_state.get((i,)).value.charge[name] for i in range(n)], dtype=np.float64)

    def occupancy(self, max_occupants):
        """(occupants[n_cells][max_occupants], surplus list) of the first internal state, in flat cell order.
        NOTE: the internal state is updated at the top of the NEXT iteration; call between iterations only
        through run_events(), which snapshots after the activator ran."""
        ist = self.mediator._activator._internal_states[0]
        cells = list(ist.cells.yield_cells())
        occ = np.full((len(cells), max_occupants), -1, dtype=np.int32)
        for index, cell in enumerate(cells):
            for s, identifier in enumerate(ist._occupants[cell]):
                occ[index, s] = self.leaf_id(identifier)
        surplus = [self.leaf_id(identifier) for identifier in ist.yield_surplus()]
        return occ, np.array(surplus, dtype=np.int32)

    def run(self, max_events=None, snapshot_every=None, max_occupants=1):
        """Run mediator.run() until EndOfRun or until max_events device events were committed. Returns the
        records. With snapshot_every = k, the state (positions, occupancy, surplus, lifting) is stored when the
        activator has just updated its internal state at a multiple of k committed events."""
        from jellyfysh.base.exceptions import EndOfRun

        class _Stop(Exception):
            pass

        act = self.mediator._activator
        run = self
        self.snapshots = []
        last_snapshot = [-1]
        # TagActivator rebinds get_event_handlers_to_run to this method after the start-of-run iteration
        # (tag_activator.py:226); an instance attribute set beforehand is what gets bound.
        orig_update = act._get_event_handlers_to_run_update

        def update(active_state, previous):
            if max_events is not None and run.events >= max_events:
                raise _Stop()
            out = orig_update(active_state, previous)
            if (snapshot_every and run.events % snapshot_every == 0 and last_snapshot[0] != run.events
                    and run._kind_of(previous) != HOST_EVENT):
                last_snapshot[0] = run.events
                run._snapshot(max_occupants)
            return out

        act._get_event_handlers_to_run_update = update
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                self.mediator.run()
        except EndOfRun:
            pass
        except _Stop:
            pass
        return np.array(self.records, dtype=RECORD_DTYPE)

    def _snapshot(self, max_occupants):
        if self.mediator._activator._internal_states:
            occ, surplus = self.occupancy(max_occupants)
        else:  # no cell system
            occ, surplus = np.full((1, max_occupants), -1, dtype=np.int32), np.zeros(0, dtype=np.int32)
        active, direction, _, stamp = self._active()
        self.snapshots.append({"event": self.events, "positions": self.positions(), "roots": self.roots(),
                               "occupants": occ,
                               "surplus": surplus, "active": active, "direction": direction,
                               "time_q": stamp.quotient, "time_r": stamp.remainder, "mode": self.mode,
                               "velocities": self.velocities()})

    def velocities(self):
        """Velocity vectors of the active units, leaf first then its ancestors, padded to three components (general
        velocities: the sequential-direction end-of-chain handler rotates them)."""
        sh = self.mediator._state_handler
        ids = sorted(sh._lifting_state._lifting_dictionary.keys(), key=len, reverse=True)
        out = np.zeros((2, 3))
        for row, identifier in enumerate(ids[:2]):
            velocity, _ = sh._lifting_state.get(identifier)
            out[row, :len(velocity)] = velocity
        return out


def default_ref_root():
    here = os.path.dirname(os.path.abspath(__file__))
    return os.environ.get("JF_REF", os.path.join(here, "..", "..", "baseline", "_ref"))
