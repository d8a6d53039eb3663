/*
 * Prints, with the reference's own hash library, the values tests/golden/ring_keys.json holds for 16 synthetic endpoints
 * (hostname "10.0.0.0", ports 1000..1015; node ids splitmix64(2i), splitmix64(2i+1)): the K=10 ring keys of every endpoint,
 * the ring-0 order and the configuration id.  It pins the oracle's reading of net.openhft:zero-allocation-hashing:0.8.
 *
 *   javac -cp zero-allocation-hashing-0.8.jar PinRingHash.java
 *   java -cp .:zero-allocation-hashing-0.8.jar PinRingHash > mine.txt     # then diff against the json's arrays
 *
 * NOT compiled in this repository's image (no JDK, no jar) — which is exactly why DESIGN.md says "ring-hash parity unpinned".
 */
import net.openhft.hashing.LongHashFunction;

import java.nio.ByteBuffer;
import java.nio.charset.StandardCharsets;
import java.util.Arrays;
import java.util.Comparator;

public final class PinRingHash {
    static long splitmix64(long x) {
        x += 0x9E3779B97F4A7C15L;
        x = (x ^ (x >>> 30)) * 0xBF58476D1CE4E5B9L;
        x = (x ^ (x >>> 27)) * 0x94D049BB133111EBL;
        return x ^ (x >>> 31);
    }

    public static void main(final String[] args) {
        final int K = 10, N = 16;
        final byte[] host = "10.0.0.0".getBytes(StandardCharsets.UTF_8);
        final long[][] key = new long[K][N];
        for (int k = 0; k < K; k++) {
            final LongHashFunction xx = LongHashFunction.xx(k);
            for (int i = 0; i < N; i++) {                 // AddressComparator.computeHash, MembershipView.java:579-582
                key[k][i] = xx.hashBytes(ByteBuffer.wrap(host)) * 31 + xx.hashInt(1000 + i);
            }
            System.out.println("keys[" + k + "] = " + Arrays.toString(key[k]));
        }
        final Integer[] ring0 = new Integer[N];
        for (int i = 0; i < N; i++) {
            ring0[i] = i;
        }
        Arrays.sort(ring0, Comparator.comparingLong(i -> key[0][i]));      // signed Long.compare
        System.out.println("ring0 = " + Arrays.toString(ring0));
        // Configuration.getConfigurationId, MembershipView.java:544-556: identifiers sorted by signed (high, low), then ring 0
        final long[][] ids = new long[N][2];
        for (int i = 0; i < N; i++) {
            ids[i][0] = splitmix64(2L * i);
            ids[i][1] = splitmix64(2L * i + 1);
        }
        Arrays.sort(ids, (a, b) -> a[0] != b[0] ? Long.compare(a[0], b[0]) : Long.compare(a[1], b[1]));
        final LongHashFunction xx0 = LongHashFunction.xx(0);
        long h = 1;
        for (final long[] id : ids) {
            h = h * 37 + xx0.hashLong(id[0]);
            h = h * 37 + xx0.hashLong(id[1]);
        }
        for (final int i : ring0) {
            h = h * 37 + xx0.hashBytes(ByteBuffer.wrap(host));
            h = h * 37 + xx0.hashInt(1000 + i);
        }
        System.out.println("configuration_id = " + h);
    }
}
