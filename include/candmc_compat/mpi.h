/* Put this directory on the include path (-Iinclude/candmc_compat) to let sources that say `#include <mpi.h>` or
 * `#include "mpi.h"` — the reference's test/MM and bench/MM mains — compile against candmc_b200's MPI subset. */
#include "../candmc/mpi.h"
