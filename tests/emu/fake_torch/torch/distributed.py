"""torch.distributed stand-in for world_size 1 (see torch/__init__.py in this directory)."""


class ReduceOp:
    MAX = "max"


def init_process_group(*a, **k):
    pass  # every "rank" of an emulated run is a lone process: collectives are identities (enough to walk the sharded code paths)


def barrier():
    pass


def all_reduce(t, op=None):
    pass


def destroy_process_group():
    pass
