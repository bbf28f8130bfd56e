// TEST / BENCH INFRASTRUCTURE (oracle/): a driver around the REFERENCE's own scalar engine, compiled by oracle/build_ref.py
// together with the reference's unmodified sources where they lie (v0/src/game/game_state.cpp, v0/src/rules/rule_engine.cpp,
// v0/src/moves/move_generator.cpp) into oracle/_ref/ref_playout.  It plays uniform-random games --
// v0::GenerateAllLegalMoves + v0::ApplyMove per ply (move_generator.hpp:54,62), at most 512 plies per game -- on N threads
// for a fixed wall-clock budget and prints one JSON line.  This is the CPU arm of BASELINE configs[1] with
// cpu_baseline.kind = "reference"; nothing in liuzhou_b200/ uses it.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "v0/game_state.hpp"
#include "v0/move_generator.hpp"

static inline uint64_t splitmix64(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    const int threads = argc > 1 ? std::atoi(argv[1]) : 1;
    const double seconds = argc > 2 ? std::atof(argv[2]) : 5.0;
    const uint64_t seed = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 20260314ull;
    std::atomic<long long> plies{0}, games{0};
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t] {
            uint64_t rng = seed * 10007ull + (uint64_t)(t + 1) * 9973ull;
            long long my_plies = 0, my_games = 0;
            while (elapsed() < seconds) {
                v0::GameState state;
                for (int ply = 0; ply < 512; ++ply) {
                    const std::vector<v0::MoveRecord> moves = v0::GenerateAllLegalMoves(state);
                    if (moves.empty()) break;
                    state = v0::ApplyMove(state, moves[splitmix64(rng) % moves.size()], true);
                    ++my_plies;
                }
                ++my_games;
            }
            plies += my_plies;
            games += my_games;
        });
    }
    for (auto& th : pool) th.join();
    const double dt = elapsed();
    std::printf("{\"plies\": %lld, \"games\": %lld, \"seconds\": %.4f, \"threads\": %d, \"plies_per_sec\": %.1f}\n",
                plies.load(), games.load(), dt, threads, plies.load() / dt);
    return 0;
}
