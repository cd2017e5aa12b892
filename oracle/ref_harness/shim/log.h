// Empty stand-in for the proprietary <log.h> that src/utils/timer.h:3 includes
// (pulled in by src/dcgrid/fluid_simulation_dcgrid.cu:4).  Nothing from it is used
// on the solver path.
#pragma once
