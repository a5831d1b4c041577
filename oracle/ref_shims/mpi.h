/* TEST INFRASTRUCTURE (oracle/): minimal stand-in so the reference's logger.hpp / profiling.hpp parse without MPI. */
#pragma once
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
