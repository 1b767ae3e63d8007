// open_chisel/weighting/ConstantWeighter.h -- the reference's header name; the class lives in b200/IntegratorPolicies.h with the other policy
// objects of the integrator.
#pragma once
#include <open_chisel/b200/IntegratorPolicies.h>
