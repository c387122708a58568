// what `configure` generates from Config.hpp.in for the non-NUMA build; the JIT paths are unused here
#ifndef SPARSEX_INTERNALS_CONFIG_HPP
#define SPARSEX_INTERNALS_CONFIG_HPP
#define SPX_JIT_INCLUDE ""
#define SPX_MULT_TEMPLATE_DIR ""
#define SPX_USE_NUMA 0
#define CLANG_INC_SEARCH_PATH ""
#endif
