"""CudaStateHandler -- the reference's state-handler contract over the chains of a device engine.

A TreeStateHandler (jellyfysh/state_handler/tree_state_handler.py) whose tree of units can be (re)filled from ONE chain of
a `jellyfysh_b200.engine.Engine`: `select_chain(c)` downloads that chain (`ecmc_download_chain`: leaf positions, root
positions, lifting state) and writes it through the contract's own `insert_into_global_state`, after which the four
methods output handlers and dumps use -- `extract_from_global_state`, `insert_into_global_state`,
`extract_active_global_state`, `extract_global_state` (state_handler.py:63-165) -- describe that chain: positions of all
units, velocity and time stamp of the active leaf unit and, for composite point objects, of its root unit (velocity x
weight, event_handler/abstracts/abstracts.py:165-190). Selected with

    [CudaBatchedMediator]
    state_handler = cuda_state_handler

    [CudaStateHandler]
    physical_state = tree_physical_state
    lifting_state = tree_lifting_state

(`jellyfysh_b200.install()` registers the module as `jellyfysh.state_handler.cuda_state_handler`). CudaBatchedMediator
binds its engines; with a plain `tree_state_handler` it fills the tree itself, chain after chain.
"""
from jellyfysh.base.time import Time
from jellyfysh.state_handler.lifting_state.tree_lifting_state import TreeLiftingState
from jellyfysh.state_handler.physical_state.tree_physical_state import TreePhysicalState
from jellyfysh.state_handler.tree_state_handler import TreeStateHandler


class CudaStateHandler(TreeStateHandler):
    """TreeStateHandler that reads its global state from a chain of a device engine."""

    def __init__(self, physical_state: TreePhysicalState, lifting_state: TreeLiftingState) -> None:
        super().__init__(physical_state, lifting_state)
        self._engines, self._shards = [], []
        self._speed, self._dimension, self._nodes_per_root = 1.0, 3, 1
        self._selected = None
        self._general = False

    def bind(self, engines, shards, speed, dimension, nodes_per_root, general_velocities=False):
        """engines with their (first chain, number of chains) blocks, in chain order."""
        self._engines, self._shards = list(engines), list(shards)
        self._speed, self._dimension, self._nodes_per_root = float(speed), int(dimension), int(nodes_per_root)
        self._general = bool(general_velocities)

    @property
    def number_of_chains(self):
        return sum(count for _, count in self._shards)

    @property
    def selected_chain(self):
        return self._selected

    def select_chain(self, chain: int) -> None:
        """Make the tree describe chain `chain`: one D2H of that chain, written through insert_into_global_state."""
        for eng, (first, count) in zip(self._engines, self._shards):
            if first <= chain < first + count:
                positions, roots, state = eng.download_chain(chain - first)
                break
        else:
            raise IndexError("chain {0} of {1}".format(chain, self.number_of_chains))
        self.load(positions, roots, state)
        self._selected = chain

    def load(self, positions, roots, state):
        """Fill the tree from arrays: positions[N][D], roots[N / nodes_per_root][D] or None, an EcmcChainState record."""
        cnodes = self.extract_global_state()
        fill_tree(cnodes, positions, roots, state, self._speed, self._dimension, self._nodes_per_root, self._general)
        self.insert_into_global_state(cnodes)


def fill_tree(cnodes, positions, roots, state, speed, dimension, nodes_per_root, general_velocities=False):
    """Write one chain's device state into root cnodes (tree_state_handler.py:213-230): positions of all units, velocity
    and time stamp of the active leaf unit and, for composite point objects, of its root unit -- velocity x weight
    (event_handler/abstracts/abstracts.py:165-190), or, for programs with general velocities, the velocities the chain
    state carries (EcmcChainState.velocity / root_velocity)."""
    npr = nodes_per_root
    active, direction = int(state["active"]), int(state["direction"])
    stamp = Time(float(state["time_q"]), float(state["time_r"]))

    def fill(unit, position, velocity):
        unit.position = [float(x) for x in position]
        unit.velocity, unit.time_stamp = velocity, (None if velocity is None else stamp)

    def along_direction(unit_speed):
        return [unit_speed if d == direction else 0.0 for d in range(dimension)]

    for index, cnode in enumerate(cnodes):
        if roots is None:
            fill(cnode.value, positions[index], along_direction(speed) if index == active else None)
            continue
        if general_velocities:
            leaf_velocity = [float(v) for v in state["velocity"][:dimension]]
            root_velocity = [float(v) for v in state["root_velocity"][:dimension]]
        else:
            leaf_velocity, root_velocity = along_direction(speed), along_direction(speed * cnode.children[0].weight)
        # EcmcChainState.mode = 1 (root-unit-active mode, dipoles/dipole_motion.ini): the root unit of the object is the
        # independent active unit, it and every one of its leaves move with the full velocity
        whole_object = int(state["mode"]) == 1 and index == active // npr
        if whole_object:
            root_velocity = along_direction(speed)
        fill(cnode.value, roots[index], root_velocity if index == active // npr else None)
        for k, child in enumerate(cnode.children):
            fill(child.value, positions[index * npr + k],
                 list(leaf_velocity) if whole_object or index * npr + k == active else None)
