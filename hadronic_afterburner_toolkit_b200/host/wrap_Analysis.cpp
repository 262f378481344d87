// Wrapper TU: the reference's Analysis driver (src/Analysis.cpp, which holds
// Analysis::HBTAnalysis, :817-835) compiled against OUR HBT_correlation and BalanceFunction classes.
// Our header shares the reference's include guard (HBT_correlation_h), so the
// `#include "HBT_correlation.h"` inside the reference source becomes a no-op.
// REF_SRC is given by the Makefile (-DREF_SRC=/root/reference/src); nothing is copied.
#include "HBT_correlation.h"
// likewise the BalanceFunction operator (Analysis::BalanceFunctionAnalysis, :856-874): guard BALANCEFUNCTION_H_
#include "BalanceFunction.h"

#define HBT_STR2(x) #x
#define HBT_STR(x) HBT_STR2(x)
#include HBT_STR(REF_SRC/Analysis.cpp)
