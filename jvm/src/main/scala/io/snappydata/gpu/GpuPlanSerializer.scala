/*
 * Serialises the matched plan fragment into the off-heap sd_plan_desc block the C ABI takes, and moves
 * buffers/rows across the JNI boundary.  NOT COMPILED in this repository's container (no JDK/Scala); a sketch
 * against the reference's class names for a maintainer to adapt (see INTEGRATION.md).
 */
package io.snappydata.gpu

import java.nio.ByteBuffer

import org.apache.spark.sql.catalyst.InternalRow
import org.apache.spark.sql.catalyst.expressions._
import org.apache.spark.sql.catalyst.expressions.aggregate._
import org.apache.spark.sql.execution.{FilterExec, ProjectExec, SparkPlan}
import org.apache.spark.sql.execution.aggregate.SnappyHashAggregateExec
import org.apache.spark.sql.execution.columnar.{ColumnBatchIterator, ColumnTableScan}
import org.apache.spark.sql.execution.metric.SQLMetric
import org.apache.spark.sql.types._
import org.apache.spark.unsafe.Platform

/** Off-heap image of sd_plan_desc (include/snappy_gpu.h) plus what the operator needs at run time. */
final class GpuPlanDesc(val address: Long, val numLiterals: Int, val scan: ColumnTableScan,
    val literals: Seq[Expression /* ParamLiteral | TokenLiteral | Literal */], val outputSchema: StructType) {
  /** evaluates the tokenised constants of THIS execution (ParamLiteral.value) into sd_literal[] */
  def literalsAddress(): Long = GpuPlanSerializer.writeLiterals(literals)
}

object GpuPlanSerializer {
  def numDevices: Int = java.lang.Integer.getInteger("snappydata.gpu.devices", 1)

  private def sdType(dt: DataType): Option[Int] = dt match {
    case BooleanType => Some(1); case ByteType => Some(2); case ShortType => Some(3); case IntegerType => Some(4)
    case LongType => Some(5); case FloatType => Some(6); case DoubleType => Some(7); case DateType => Some(8)
    case TimestampType => Some(9); case StringType => Some(10)
    case d: DecimalType if d.precision <= 18 => Some(11)
    case _ => None
  }

  /** Some(desc) iff the fragment is SnappyHashAggregateExec(Partial) over [Project] [Filter] ColumnTableScan with
    * expressions the C ABI can express (sd_op in include/snappy_gpu.h); anything else stays on the stock operators. */
  def tryBuild(agg: SnappyHashAggregateExec): Option[GpuPlanDesc] = {
    def unwrap(p: SparkPlan, filters: Seq[Expression]): Option[(ColumnTableScan, Seq[Expression])] = p match {
      case FilterExec(cond, child) => unwrap(child, filters :+ cond)
      case ProjectExec(_, child) => unwrap(child, filters) // projections are inlined into the expression trees
      case scan: ColumnTableScan => Some((scan, filters))
      case _ => None
    }
    unwrap(agg.child, Nil).flatMap { case (scan, filters) =>
      // flatten: scan.output -> sd_column[]; filters.reduce(And) / grouping / aggregate children -> sd_expr[]
      // (Cast(l_shipdate, StringType) >= '1994-01-01' produced by PromoteStrings is normalised to an int-day compare,
      //  SURVEY.md Appendix B.8); literals become slots in evaluation order
      ??? // elided: mechanical tree walk writing the structs of include/snappy_gpu.h with Platform.putInt/putLong
    }
  }

  def writeLiterals(literals: Seq[Expression]): Long = ???

  /** per batch: value buffers via getColumnLob, deltas via the iterator's delta lookups (ColumnBatchIterator.scala:122-163) */
  def submitBatch(plan: Long, it: ColumnBatchIterator, stats: ByteBuffer, desc: GpuPlanDesc): Unit = ???

  /** un-rolled-over rows as UnsafeRows (ColumnTableScan.scala:572-588) */
  def submitRowBuffer(plan: Long, rows: Iterator[InternalRow], desc: GpuPlanDesc): Unit = ???

  /** sd_plan_finish -> Iterator[UnsafeRow] over the returned [int64 size][row] stream (UnsafeRow.pointTo) */
  def finishToUnsafeRows(plan: Long, desc: GpuPlanDesc, numOutputRows: SQLMetric): Iterator[InternalRow] = ???
}
