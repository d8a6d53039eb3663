/*
 * Read side of com.vrg.rapid.MembershipView served from the device-resident K-ring view.  UNCOMPILED here (no JDK).
 * Built once per configuration from the full endpoint list (ring mutations stay in the Java view, which the service
 * keeps for the join protocol); ids: members 0..n-1 in list order, joiners registered afterwards.
 */
package com.vrg.rapid;

import com.vrg.rapid.gpu.Native;
import com.vrg.rapid.pb.Endpoint;

import java.io.ByteArrayOutputStream;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

final class GpuMembershipView {
    private final int K;
    private final long handle;
    private final List<Endpoint> byId = new ArrayList<>();
    private final Map<Endpoint, Integer> ids = new HashMap<>();
    private final int members;

    GpuMembershipView(final int K, final List<Endpoint> endpoints, final int device) {
        this.K = K;
        final ByteArrayOutputStream bytes = new ByteArrayOutputStream();
        final int[] off = new int[endpoints.size() + 1];
        final int[] port = new int[endpoints.size()];
        int i = 0;
        for (final Endpoint e : endpoints) {
            final byte[] h = e.getHostname().toByteArray();
            bytes.write(h, 0, h.length);
            off[i + 1] = off[i] + h.length;
            port[i] = e.getPort();
            ids.put(e, i);
            byId.add(e);
            i++;
        }
        this.members = endpoints.size();
        this.handle = Native.viewCreate(K, members, bytes.toByteArray(), off, port, device);
        if (handle == 0) {
            throw new IllegalStateException(Native.lastError());
        }
    }

    long handle() {
        return handle;
    }

    boolean isHostPresent(final Endpoint e) {                 // MembershipView.java:330-337
        final Integer id = ids.get(e);
        return id != null && id < members;
    }

    int idOf(final Endpoint e, final boolean registerAsJoiner) {
        Integer id = ids.get(e);
        if (id == null) {
            if (!registerAsJoiner) {
                throw new MembershipView.NodeNotInRingException(e);
            }
            final byte[] h = e.getHostname().toByteArray();
            id = Native.viewRegisterJoiners(handle, h, new int[]{0, h.length}, new int[]{e.getPort()});
            ids.put(e, id);
            byId.add(e);
        }
        return id;
    }

    Endpoint endpointOf(final int id) {
        return byId.get(id);
    }

    List<Endpoint> getObserversOf(final Endpoint node) {      // :210-224
        return row(node, true);
    }

    List<Endpoint> getSubjectsOf(final Endpoint node) {       // :267-282
        return row(node, false);
    }

    List<Endpoint> getExpectedObserversOf(final Endpoint node) {   // :292-303
        final int[] out = new int[K];
        final int n = Native.viewExpectedObservers(handle, node.getHostname().toByteArray(), node.getPort(), out);
        return map(out, n);
    }

    private List<Endpoint> row(final Endpoint node, final boolean observers) {
        final Integer id = ids.get(node);
        if (id == null || id >= members) {
            throw new MembershipView.NodeNotInRingException(node);
        }
        final int[] out = new int[K];
        final int n = observers ? Native.viewObservers(handle, id, out) : Native.viewSubjects(handle, id, out);
        return map(out, n);
    }

    private List<Endpoint> map(final int[] out, final int n) {
        if (n < 0) {
            throw new IllegalStateException(Native.lastError());
        }
        final List<Endpoint> l = new ArrayList<>(n);
        for (int i = 0; i < n; i++) {
            l.add(byId.get(out[i]));
        }
        return l;
    }
}
