// YAML run report (ref/output.cpp:48-494): run configuration, the thermo records with 10
// significant digits, the time split and the per-rank count histograms.
#pragma once
class Simulation;
void output(Simulation& sim);
