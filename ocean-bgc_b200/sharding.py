"""Column-slab sharding of a mesh over ranks (one process per GPU).

Columns are independent (every loop of the reference indexes (..., column, ...)
only: BGC_mod.F90:733/799/2808, DMS_mod.F90:459/495/852, MACROS_mod.F90:301/331),
so rank r simply owns a contiguous slab; there is no halo and no data-path
collective.  The only exchange is the all-reduce of the 64-double inventory
vector (bgc_inventory_allreduce).
"""


def slab(rank, world, n_columns):
    """(first_column, n_local) of rank's contiguous slab; slabs differ by at most one column."""
    if world < 1 or not (0 <= rank < world) or n_columns < 0:
        raise ValueError("bad slab request rank=%r world=%r n=%r" % (rank, world, n_columns))
    base, rem = divmod(n_columns, world)
    n_local = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n_local


def slabs(world, n_columns):
    return [slab(r, world, n_columns) for r in range(world)]
