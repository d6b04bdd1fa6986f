//  Umbrella header (reference: graph_framework/graph_framework.hpp).
#ifndef gfb_graph_framework_hpp
#define gfb_graph_framework_hpp
#include "node.hpp"
#include "vector.hpp"
#include "emit.hpp"
#include "jit.hpp"
#include "workflow.hpp"
#include "newton.hpp"
#include "equilibrium.hpp"
#include "dispersion.hpp"
#include "solver.hpp"
#include "absorption.hpp"
#endif
