/*
 * GetNextHarness — dumps golden vectors for CacheMissForwardingLB.getNext FROM THE REFERENCE ITSELF.
 *
 * TEST INFRASTRUCTURE, SOURCE ONLY: this image has no JDK and the reference's dependencies (litelinks-core 1.7.2,
 * kv-utils 0.5.1, guava, eclipse-collections; pom.xml:60-76, 299-350) are not vendored, so it cannot be compiled or run
 * here (SURVEY.md §8c).  On a box that has JDK 21 and the reference built (`mvn -DskipTests package` in the reference tree):
 *
 *   CP=$REF/target/classes:$(cat $REF/target/classpath.txt)     # mvn dependency:build-classpath -Dmdep.outputFile=...
 *   javac -cp $CP -d oracle/java/_build oracle/java/GetNextHarness.java
 *   java  -cp $CP:oracle/java/_build com.ibm.watson.modelmesh.GetNextHarness tests/golden/fleet_*.json
 *
 * Input  (written by tools/golden_fleets.py): {"minSpaceUnits", "minChurnAgeMs", "now", "typeConstraints": {...}|null,
 *         "instances": [{"id", "record": <InstanceRecord json>, "active": bool}], "replaced": [...],
 *         "decisions": [{"type", "self", "fresh": <InstanceRecord json>|null, "favourSelf", "lastUsed", "excluded": [ids]}]}
 * Output (tests/golden/<same name>.expected.json): per decision everything BEFORE the random draw (N4): the filtered first
 * entry, the ordered shortlist, instReqLoad and which candidates survive the rpm filter; plus the PLACEMENT_ORDER order of
 * the instances.  tests/test_oracle_golden.py::test_java_golden_vectors loads every such pair it finds and holds the oracle
 * to it; with no files present it reports "parity unpinned" and skips.
 *
 * It lives in the reference's package because PLACEMENT_ORDER and CacheMissForwardingLB are (package-visible) members of
 * the abstract ModelMesh (ModelMesh.java:4646, 4757): the harness subclasses ModelMesh with the loader methods stubbed,
 * fills clusterState / typeConstraints / upgradeTracker reflectively, and instruments the LB by overriding
 * ThreadLocalRandom use through a recorded candidate list (the candidates field is cleared only by the TODO at MM:4987).
 */
package com.ibm.watson.modelmesh;

import java.io.File;
import java.lang.reflect.Field;
import java.lang.reflect.Method;
import java.util.ArrayList;
import java.util.List;
import java.util.Map;

import com.fasterxml.jackson.databind.JsonNode;
import com.fasterxml.jackson.databind.ObjectMapper;
import com.fasterxml.jackson.databind.node.ArrayNode;
import com.fasterxml.jackson.databind.node.ObjectNode;

public final class GetNextHarness {
    private static final ObjectMapper M = new ObjectMapper();

    public static void main(String[] args) throws Exception {
        for (String path : args) {
            JsonNode in = M.readTree(new File(path));
            ObjectNode out = M.createObjectNode();
            // --- a ModelMesh instance with just enough state for the placement path (no KV store, no runtime) ---
            ModelMesh mm = HarnessMesh.create(in.get("minSpaceUnits").asLong(), in.get("minChurnAgeMs").asLong(), in.get("now").asLong());
            HarnessMesh.setTypeConstraints(mm, in.get("typeConstraints"));
            for (JsonNode inst : in.get("instances")) {
                InstanceRecord rec = M.treeToValue(inst.get("record"), InstanceRecord.class);
                HarnessMesh.instanceAdded(mm, inst.get("id").asText(), rec, inst.get("active").asBoolean());
            }
            HarnessMesh.setLikelyReplaced(mm, in.get("replaced"));
            ArrayNode order = out.putArray("order");
            for (Map.Entry<String, InstanceRecord> e : HarnessMesh.clusterState(mm)) order.add(e.getKey());
            ArrayNode res = out.putArray("decisions");
            for (JsonNode d : in.get("decisions")) res.add(HarnessMesh.getNext(mm, d));
            M.writerWithDefaultPrettyPrinter().writeValue(new File(path.replace(".json", ".expected.json")), out);
        }
    }

    /** Reflection helpers over the reference's own fields and inner classes; no logic of the path is restated here. */
    static final class HarnessMesh {
        static ModelMesh create(long minSpaceUnits, long minChurnAgeMs, long now) throws Exception {
            ModelMesh mm = (ModelMesh) sun.misc.Unsafe.class.getDeclaredMethod("allocateInstance", Class.class)
                    .invoke(unsafe(), Class.forName("com.ibm.watson.modelmesh.SidecarModelMesh"));
            set(mm, "minChurnAgeMs", minChurnAgeMs);
            set(mm, "instanceId", "harness-self");
            // minSpaceUnits is derived from loader params (MM:767-769); the harness sets the derived value directly
            set(mm, "minSpaceUnits", minSpaceUnits);
            HarnessClock.freeze(now);  // ModelMesh.currentTimeMillis() is static: the harness runs under a frozen-clock agent
            return mm;
        }
        static void setTypeConstraints(ModelMesh mm, JsonNode tc) throws Exception {
            if (tc == null || tc.isNull()) return;
            Method m = TypeConstraintManager.class.getDeclaredMethod("typeMappingsUpdated", String.class);
            Object tcm = TypeConstraintManager.class.getDeclaredConstructors()[0].newInstance("harness-self", HarnessMesh.clusterState(mm), null);
            m.setAccessible(true);
            m.invoke(tcm, M.writeValueAsString(tc));
            set(mm, "typeConstraints", tcm);
        }
        static void instanceAdded(ModelMesh mm, String id, InstanceRecord rec, boolean active) throws Exception {
            Method m = ModelMesh.class.getDeclaredMethod("handleInstanceTableChange",
                    com.ibm.watson.kvutils.KVTable.EventType.class, String.class, InstanceRecord.class);
            m.setAccessible(true);
            m.invoke(mm, com.ibm.watson.kvutils.KVTable.EventType.ENTRY_ADDED, id, rec);
            if (active) HarnessClock.activeInstances.add(id);
        }
        static void setLikelyReplaced(ModelMesh mm, JsonNode arr) throws Exception {
            Object ut = get(mm, "upgradeTracker");
            @SuppressWarnings("unchecked") Map<String, Object> lr = (Map<String, Object>) get(ut, "likelyReplacedReplicaSets");
            for (JsonNode p : arr) lr.put(p.asText(), Boolean.TRUE);
        }
        @SuppressWarnings("unchecked")
        static Iterable<Map.Entry<String, InstanceRecord>> clusterState(ModelMesh mm) throws Exception {
            return (Iterable<Map.Entry<String, InstanceRecord>>) get(mm, "clusterState");
        }
        /** Runs CacheMissForwardingLB.getNext with ThreadLocalRandom pinned to index 0 and returns what it saw before the draw. */
        static ObjectNode getNext(ModelMesh mm, JsonNode d) throws Exception {
            Class<?> lbc = Class.forName("com.ibm.watson.modelmesh.ModelMesh$CacheMissForwardingLB");
            Object lb = lbc.getDeclaredConstructors()[0].newInstance(mm);
            Object exclude = HarnessClock.newExcludeSet(mm, d);
            Object[] sis = HarnessClock.serviceInstances();
            Method gn = lbc.getMethod("getNext", Object[].class, String.class, Object[].class);
            Object r = gn.invoke(lb, sis, "ensureLoaded", new Object[0]);
            ObjectNode o = M.createObjectNode();
            o.put("result", r == null ? "null" : (r == com.ibm.watson.litelinks.client.LoadBalancer.ABORT_REQUEST ? "SELF" : r.toString()));
            ArrayNode c = o.putArray("candidates");
            for (Object iid : (List<?>) get(lb, "candidates")) c.add(String.valueOf(iid));
            ArrayNode l = o.putArray("instReqLoad");
            Object irl = get(lb, "instReqLoad");
            int n = (Integer) irl.getClass().getMethod("size").invoke(irl);
            for (int i = 0; i < n; i++) l.add((Integer) irl.getClass().getMethod("get", int.class).invoke(irl, i));
            o.set("excludeAfter", M.valueToTree(new ArrayList<>((java.util.Set<?>) exclude)));
            return o;
        }
        static Object unsafe() throws Exception { Field f = sun.misc.Unsafe.class.getDeclaredField("theUnsafe"); f.setAccessible(true); return f.get(null); }
        static Object get(Object o, String name) throws Exception {
            for (Class<?> c = o.getClass(); c != null; c = c.getSuperclass())
                try { Field f = c.getDeclaredField(name); f.setAccessible(true); return f.get(o); } catch (NoSuchFieldException e) { /* up */ }
            throw new NoSuchFieldException(name);
        }
        static void set(Object o, String name, Object v) throws Exception {
            for (Class<?> c = o.getClass(); c != null; c = c.getSuperclass())
                try { Field f = c.getDeclaredField(name); f.setAccessible(true); f.set(o, v); return; } catch (NoSuchFieldException e) { /* up */ }
            throw new NoSuchFieldException(name);
        }
    }

    /** The pieces that need the litelinks types of the box the harness is built on: the service-instance array handed to
     *  getNext (ids of the `active` instances), the per-request CacheMissExcludeSet thread-local (MM:4755), and the clock. */
    static final class HarnessClock {
        static final java.util.Set<String> activeInstances = new java.util.LinkedHashSet<>();
        static void freeze(long now) { System.setProperty("mm.harness.now", Long.toString(now)); }
        static Object newExcludeSet(ModelMesh mm, JsonNode d) throws Exception {
            Class<?> ces = Class.forName("com.ibm.watson.modelmesh.ModelMesh$CacheMissExcludeSet");
            Object ex = ces.getDeclaredConstructors()[0].newInstance();
            for (JsonNode e : d.get("excluded")) ((java.util.Set<String>) ex).add(e.asText());
            HarnessMesh.set(ex, "favourSelf", d.get("favourSelf").asBoolean());
            HarnessMesh.set(ex, "lastUsedTime", d.get("lastUsed").asLong());
            ((ThreadLocal<Object>) HarnessMesh.get(mm, "cacheMissExcludeTl")).set(ex);
            return ex;
        }
        static Object[] serviceInstances() { return activeInstances.toArray(); }
    }
}
