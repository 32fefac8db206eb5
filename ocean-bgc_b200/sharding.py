"""Column-slab sharding of a mesh over ranks (one process per GPU).

Columns are independent (every loop of the reference indexes (..., column, ...)
only: BGC_mod.F90:733/799/2808, DMS_mod.F90:459/495/852, MACROS_mod.F90:301/331),
so rank r simply owns a contiguous slab; there is no halo and no data-path
collective.  The only exchange is the all-reduce of the 64-double inventory
vector (bgc_inventory_allreduce).
"""


def slab(rank, world, n_columns, even=False):
    """(first_column, n_local) of rank's contiguous slab.  By default the slabs differ by at most
    one column.  even=True makes every slab but the last an even number of columns
    (ceil(n/world) rounded up to even): with an even numColumnsMax every level row of the device
    arrays stays 16-byte aligned, which the column sweep needs for its bulk (TMA) copies."""
    if world < 1 or not (0 <= rank < world) or n_columns < 0:
        raise ValueError("bad slab request rank=%r world=%r n=%r" % (rank, world, n_columns))
    if even:
        per = -(-n_columns // world)
        per += per & 1
        first = min(rank * per, n_columns)
        return first, max(0, min(per, n_columns - first))
    base, rem = divmod(n_columns, world)
    n_local = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n_local


def slabs(world, n_columns, even=False):
    return [slab(r, world, n_columns, even) for r in range(world)]
