// placeholder main until the CLI lands (next commit)
#include <cstdio>
int main() { std::puts("bronko (b200) CLI: not built yet"); return 1; }
