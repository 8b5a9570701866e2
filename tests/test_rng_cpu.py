"""iss_rng.h (shared by the CUDA kernels, the C oracle and the host): Stream::skip(n) leaves the stream
where n calls of next() would leave it (the count pass of the decay kernel relies on it)."""
import os
import subprocess
import textwrap

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = textwrap.dedent("""
    #include <cstdio>
    #include <cstdint>
    #include <cstring>
    #include "iss_rng.h"
    using namespace iss;
    int main() {
        int bad = 0, checked = 0;
        for (int pre = 0; pre < 9; pre++)          // words consumed before, odd counts included
            for (int n = 0; n < 14; n++) {
                Stream a, b;
                a.init(12345, 3, 7, 11, 13);
                b.init(12345, 3, 7, 11, 13);
                for (int i = 0; i < pre; i++) { a.word(); b.word(); }
                for (int i = 0; i < n; i++) a.next();
                b.skip(n);
                for (int i = 0; i < 10; i++, checked++)
                    if (a.next() != b.next()) bad++;
            }
        printf("%d %d\\n", bad, checked);
        return 0;
    }
""")


def test_stream_skip_equals_discarded_draws(tmp_path):
    src = tmp_path/"skip.cpp"
    src.write_text(SRC)
    exe = tmp_path/"skip"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(REPO, "iss_b200", "csrc"), str(src),
                    "-o", str(exe)], check=True)
    bad, checked = map(int, subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split())
    assert checked == 9*14*10 and bad == 0
